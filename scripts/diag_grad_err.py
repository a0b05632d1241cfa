"""Diagnostic: worst per-parameter gradient errors of the bf16 generator backward vs the fp32 oracle (the quantity
tests/test_gpu_model_parity.py::test_generator_backward_parity bounds by 5e-2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sradsgan_oracle as O
from sradsgan_b200 import ops
from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup

ops.set_precision("bf16")
scale, ng, nb = 4, 2, 1
sd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=5, init="fan"))
if os.environ.get("SR_DIAG_SOFT", "1") == "1":
    for k in sd:
        if k.startswith(("conv1.0.", "MSB.")):
            sd[k] = sd[k] * 0.1
G = GeneratorResNet(ResGroup, n_residual_blocks=ng, n_basic_blocks=nb, upscale_factor=scale)
G.load_state_dict(sd, strict=True)
G.cuda()
lr, hr = O.synthetic_batch(2, scale, 64, seed=9)
y = G(lr.cuda())
(0.5 * (y.float() - hr.cuda()) ** 2).mean().backward()
mine = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
O.tie_upsampling(mine)
(0.5 * (O.generator_forward(mine, lr, scale, ng, nb) - hr) ** 2).mean().backward()
y_ref = O.generator_forward({k: v.detach() for k, v in mine.items()}, lr, scale, ng, nb)
print("forward rel err %.3e" % ((y.detach().float().cpu().double() - y_ref.double()).norm() / y_ref.double().norm()).item())
rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
errs = sorted(((rel(p.grad, mine[k].grad), k) for k, p in G.named_parameters() if k not in O.NOISE_GRAD_KEYS), reverse=True)
print("SR_LA_MMA=%s" % os.environ.get("SR_LA_MMA", "1"))
for e, k in errs[:12]:
    print("  %.4f  %s" % (e, k))
