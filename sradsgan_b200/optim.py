"""Flat-buffer Adam: all parameters of a network live in ONE contiguous fp32 buffer (and so do their
gradients and both Adam moments), so that
  * `zero_grad()` is one memset, `step()` is one fused kernel (sr_adam_step, optionally with the WGAN
    weight clamp of reference model/sradsgan.py:891-892 folded in),
  * the gradient buffer IS the all-reduce payload of data-parallel training (sradsgan_b200/dp.py).
Semantics follow torch.optim.Adam as used by the reference (model/sradsgan.py:724-725): lr 2e-4,
betas (0.9, 0.999), eps 1e-8, no weight decay, bias correction by step count."""
from collections import OrderedDict

import torch

from . import _lib, ops


class FlatAdam:
    def __init__(self, module_or_params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, clamp=None, chunk_of=None):
        if isinstance(module_or_params, torch.nn.Module):
            named = [(n, p) for n, p in module_or_params.named_parameters() if p.requires_grad]
        else:
            named = [("p%d" % i, p) for i, p in enumerate(module_or_params)]
        if not named:
            raise ValueError("FlatAdam: no trainable parameters")
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device
        total = sum((p.numel() + 3) // 4 * 4 for p in self.params)     # every parameter starts 16-byte aligned
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets = OrderedDict()
        off = 0
        for n, p in named:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view(p.shape)
            p.grad = self.flat_grad[off:off + k].view(p.shape)
            p._sr_flat_grad = p.grad       # ops.py: first-order backward passes accumulate straight into this view
            self.offsets[n] = (off, k)
            off += (k + 3) // 4 * 4
        # ops.packed(): packed bf16 operands are validated against this stamp.  Every update of the masters through raw
        # pointers (step(), a CUDA-graph replay of step()) must call touch().
        self.gen = [ops.next_generation()]
        for p in self.params:
            p._sr_genref = self.gen
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps}]
        self.clamp = clamp
        self.step_count = 0
        self.step_t = torch.zeros(1, dtype=torch.int32, device=dev)   # device-side copy (graph replay)
        # contiguous chunks (name-prefix -> [start, end)) used by the overlapped all-reduce
        self.chunks = self._make_chunks(chunk_of) if chunk_of is not None else [("all", 0, total, list(self.names))]

    def _make_chunks(self, chunk_of):
        chunks = []
        for n in self.names:
            key = chunk_of(n)
            off, k = self.offsets[n]
            if chunks and chunks[-1][0] == key:
                chunks[-1][2] = off + k
                chunks[-1][3].append(n)
            else:
                chunks.append([key, off, off + k, [n]])
        return [tuple(c) for c in chunks]

    def _rebind(self, copy):
        """autograd normally accumulates in place into the flat views; if a .grad was ever replaced
        (None, or an out-of-place accumulation), fold it back so the flat buffer stays authoritative."""
        base = self.flat_grad.data_ptr()
        for p, n in zip(self.params, self.names):
            off, k = self.offsets[n]
            if p.grad is None:
                p.grad = self.flat_grad[off:off + k].view(p.shape)
            elif p.grad.data_ptr() != base + 4 * off:
                if copy:
                    self.flat_grad[off:off + k].copy_(p.grad.detach().reshape(-1).float())
                p.grad = self.flat_grad[off:off + k].view(p.shape)

    def zero_grad(self, set_to_none=False):
        ops.wgrad_join()
        self.flat_grad.zero_()
        self._rebind(copy=False)

    def step(self, grad_scale=1.0):
        ops.wgrad_join()               # weight gradients launched on the side stream (ops._WgradStream)
        self._rebind(copy=True)
        self.step_count += 1
        self.step_t += 1
        g = self.param_groups[0]
        _lib.backend().adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0],
                                 g["betas"][1], g["eps"], self.step_count, grad_scale, self.clamp, step_tensor=self.step_t)
        self.touch()

    def touch(self):
        """the parameters changed through raw pointers (this step, or a graph replay of it): invalidate THIS network's
        cached packed operands"""
        self.gen[0] = ops.next_generation()

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [dict(g) for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        self.step_t.fill_(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups = [dict(g) for g in sd["param_groups"]]
