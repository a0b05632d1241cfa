"""GPU: the NDSRGAN sibling on the CUDA path — 64-output-channel convolutions on the tcgen05 halo kernel, growth layers
and the 4x4 critic convolutions on the library's other kernels — against the CPU oracle (oracle/ndsrgan_oracle.py) and the golden
vectors recorded from the UNMODIFIED reference `model.ndsrgan`.

Tolerances (BASELINE.json north_star): per-layer relative L2 error <= 1e-4 in fp32 mode, <= 1e-2 in bf16 mode."""
import os

import numpy as np
import pytest
import torch

from oracle import ndsrgan_oracle as N
from oracle import sradsgan_oracle as O
from test_srgan_cpu import srgan_args

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def ngolden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "ndsrgan_golden.pt"), weights_only=False)


@pytest.fixture()
def precision(request):
    from sradsgan_b200 import ops
    prev = ops.config.compute_dtype
    ops.set_precision(request.param)
    yield request.param
    ops.config.compute_dtype = prev


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
@pytest.mark.parametrize("scale", [4, 3])
def test_generator_per_block_parity(precision, scale):
    from sradsgan_b200.model.ndsrgan import GeneratorResNet
    n_blocks = 3
    sd = N.make_gen_state(scale, n_blocks, 13 + scale)
    net = GeneratorResNet(upscale_factor=scale, n_blocks=n_blocks)
    net.load_state_dict(sd, strict=True)
    net.cuda().train()
    lr, hr = N.synthetic_batch(4, scale, 24 * scale, seed=3)
    got, hooks = {}, []
    for k in range(1, n_blocks + 1):
        hooks.append(getattr(net.DCRDB_block, "DRRDB%d" % k).register_forward_hook(
            lambda m, inp, o, k="DRRDB%d" % k: got.__setitem__(k, o.detach().float().cpu())))
    with torch.no_grad():
        y = net(lr.cuda()).float().cpu()
    for h in hooks:
        h.remove()
    taps = {}
    with torch.no_grad():
        y_ref = N.generator_forward(sd, lr, scale, n_blocks, taps)
    tol = TOL[precision]
    worst = max((rel(v, taps[k]), k) for k, v in got.items())
    assert worst[0] < tol, "per-block error %g at %s" % worst
    assert rel(y, y_ref) < 2 * tol        # behind the last tap: conv2, one or two nearest + conv stages, two output convolutions


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
def test_critic_forward_and_input_gradient(precision):
    from sradsgan_b200.model.ndsrgan import Discriminator
    dsd = N.make_state(N.discriminator_spec(), seed=6, init="fan")
    D = Discriminator()
    D.load_state_dict(dsd, strict=True)
    D.cuda().train()
    x = torch.rand(8, 3, 96, 96, generator=torch.Generator().manual_seed(1))
    xg = x.clone().cuda().requires_grad_(True)
    y = D(xg)
    (y.float() ** 2).mean().backward()
    ref = {k: v.clone() for k, v in dsd.items()}
    xr = x.clone().requires_grad_(True)
    y_ref = N.discriminator_forward(ref, xr)
    (y_ref ** 2).mean().backward()
    tol = TOL[precision]
    assert y.shape == y_ref.shape and rel(y, y_ref) < 2 * tol          # five layers end to end
    # input gradient through three train-mode BatchNorm layers: fp32 mode holds the wiring to 2e-3; in bf16 mode the BatchNorm backward
    # (differences of batch statistics of bf16 activations) carries the error level the SRADSGAN critic shows at full size
    # (profiles/r02_parity_fullsize_bf16.txt: D gradients median 5e-2, worst 2e-1 against float64)
    assert rel(xg.grad, xr.grad) < (2e-3 if precision == "fp32" else 1.5e-1)
    for k in dsd:
        if "running" in k:
            assert rel(D.state_dict()[k], ref[k]) < tol, k


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
def test_trainer_steps_vs_reference_golden(precision, ngolden):
    from sradsgan_b200 import ops
    from sradsgan_b200.model.ndsrgan import NDSRGAN
    c = ngolden["train_steps"]["cfg"]
    G = N.make_gen_state(c["scale"], 23, c["gseed"])
    D = N.make_state(N.discriminator_spec(), seed=c["dseed"], init="fan")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    net = NDSRGAN(srgan_args(model_name="NDSRGAN", lr=c["lr"], scale_factor=c["scale"], batch_size=c["batch"], vgg_state=V, precision=precision))
    net.build(init=False)
    net.generator.load_state_dict(G, strict=True)
    net.discriminator.load_state_dict(D, strict=True)
    ops.bump_weight_generation()
    tol = 1e-3 if precision == "fp32" else 5e-2
    for it, want in enumerate(ngolden["train_steps"]["steps"]):
        lr, hr = N.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        out = net.train_step(lr.cuda(), hr.cuda())
        for k in ("loss_G", "loss_D", "pixel", "content", "adv"):
            assert abs(out[k].item() - want[k]) <= tol * max(1.0, abs(want[k])), (it, k, out[k].item(), want[k])


def test_graphed_step_runs():
    from sradsgan_b200 import ops
    from sradsgan_b200.model.ndsrgan import NDSRGAN
    prev = ops.config.compute_dtype
    try:
        V = O.make_state(O.vgg_spec(), seed=5, init="fan")
        lr, hr = N.synthetic_batch(4, 4, 96, seed=2)
        res = {}
        for mode in ("eager", "graph"):
            torch.manual_seed(0); np.random.seed(0)
            net = NDSRGAN(srgan_args(model_name="NDSRGAN", scale_factor=4, batch_size=4, crop_size=96, vgg_state=V, precision="bf16", seed=3))
            net.n_blocks = 2
            net.build(init=True)
            fn = net.train_step if mode == "eager" else net.graphed_step
            for _ in range(2):
                out = fn(lr.cuda(), hr.cuda())
            res[mode] = (out["loss_G"].item(), out["pixel"].item())
            assert all(np.isfinite(v) for v in res[mode])
        assert abs(res["eager"][1] - res["graph"][1]) <= 2e-2 * max(1.0, abs(res["eager"][1]))
    finally:
        ops.config.compute_dtype = prev
