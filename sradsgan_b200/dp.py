"""Data-parallel training over one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The reference is single-GPU (`device_ids=[0]`, model/sradsgan.py:659; its nn.DataParallel branch :691-694
is unreachable), so this is new functionality designed for the 8xB200 box:
  * pure data parallelism — the batch is sharded per rank, parameters/Adam state are replicated;
  * ONE sum all-reduce per network per step over the flat gradient buffer of FlatAdam, divided by the
    world size inside the fused Adam kernel (`grad_scale`), so no extra pass over the gradients;
  * the generator's bucket is cut into contiguous chunks (one per ResGroup / top-level block) and each
    chunk's all-reduce is enqueued on a side stream as soon as autograd has accumulated the last
    gradient of that chunk, overlapping communication with the rest of backward;
  * BatchNorm statistics in D stay per-rank (the reference semantics at batch 16 per GPU; no SyncBN).
"""
import torch
import torch.distributed as dist


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size():
    return dist.get_world_size() if is_dist() else 1


def all_reduce_flat(buf, group=None):
    """sum all-reduce of a flat gradient buffer on the current stream (no-op for one process)"""
    from . import ops
    ops.wgrad_join()
    if is_dist():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)


class NullReducer:
    """stands in for BucketReducer while CUDA graphs are being captured (collectives are issued between replays)"""

    def __init__(self, world=1):
        self.world = world

    def arm(self):
        pass

    def finish(self):
        return 1.0 / self.world


class BucketReducer:
    """Overlapped chunked all-reduce of a FlatAdam gradient buffer."""

    def __init__(self, opt, overlap=True, group=None):
        self.opt = opt
        self.group = group
        self.world = dist.get_world_size(group) if is_dist() else 1
        self.overlap = overlap and self.world > 1
        self.cuda = opt.flat_grad.is_cuda
        self.comm_stream = torch.cuda.Stream() if (self.cuda and self.world > 1) else None
        self.pending = {}
        self.armed = False
        self.launched = []
        self._handles = []
        if self.overlap:
            name_to_chunk = {}
            for ci, (_, s, e, names) in enumerate(opt.chunks):
                for n in names:
                    name_to_chunk[n] = ci
            self.chunk_size = [len(c[3]) for c in opt.chunks]
            for p, n in zip(opt.params, opt.names):
                ci = name_to_chunk[n]
                # fires once per backward per parameter — also when a kernel accumulated the gradient in place and
                # the Function returned None for it (the AccumulateGrad node still runs its post hooks)
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(ci)))

    def _make_hook(self, ci):
        def hook(param):
            if not self.armed:
                return
            left = self.pending.get(ci, self.chunk_size[ci]) - 1
            self.pending[ci] = left
            if left == 0:
                self._launch(ci)
        return hook

    def _launch(self, ci):
        from . import ops
        ops.wgrad_join()               # side-stream weight gradients must have landed in the bucket
        _, s, e, _ = self.opt.chunks[ci]
        buf = self.opt.flat_grad[s:e]
        if self.comm_stream is not None:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        self.launched.append(ci)

    def arm(self):
        """Call right before the backward pass whose gradients complete the bucket."""
        self.pending = {}
        self.launched = []
        self.armed = self.overlap

    def finish(self):
        """Call after backward: reduces whatever was not overlapped, then makes the compute stream wait.
        Returns the factor FlatAdam.step must apply (1/world)."""
        from . import ops
        ops.wgrad_join()
        if self.world > 1:
            if self.overlap:
                for ci in range(len(self.opt.chunks)):
                    if ci not in self.launched:
                        self._launch(ci)
            else:
                buf = self.opt.flat_grad
                if self.comm_stream is not None:
                    self.comm_stream.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(self.comm_stream):
                        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                else:
                    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.armed = False
        return 1.0 / self.world


def generator_chunk_key(name):
    """one chunk per top-level block, one per residual group"""
    parts = name.split(".")
    return ".".join(parts[:2]) if parts[0] == "res_groups" else parts[0]


def broadcast_parameters(opt, src=0, group=None):
    """replicas start from identical weights"""
    if is_dist():
        dist.broadcast(opt.flat_param, src=src, group=group)
