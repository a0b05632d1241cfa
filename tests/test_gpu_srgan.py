"""GPU: the SRGAN sibling on the CUDA path — modules -> autograd Functions -> C ABI -> the sm_100a kernels of the
SRADSGAN hot path — against the CPU oracle (oracle/srgan_oracle.py) and the golden vectors recorded from the UNMODIFIED reference
`model.srgan` classes.

Tolerances (BASELINE.json north_star): per-layer relative L2 error <= 1e-4 in fp32 mode, <= 1e-2 in bf16 mode; the train-mode
BatchNorm stack is evaluated at a batch where its statistics are well conditioned (16 x 24^2 LR pixels per channel)."""
import os

import pytest
import torch

from oracle import sradsgan_oracle as O
from oracle import srgan_oracle as S
from oracle.make_golden import summarize
from oracle.make_golden_srgan import SRGAN_CASES
from test_srgan_cpu import srgan_args

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def sgolden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "srgan_golden.pt"), weights_only=False)


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
    from sradsgan_b200.model.srgan import GeneratorResNet
    n_res = 4
    sd = S.tie_upsampling(S.make_state(S.generator_spec(scale, n_res), seed=11 + scale, init="fan"))
    net = GeneratorResNet(n_residual_blocks=n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.cuda().train()
    lr, hr = S.synthetic_batch(16, scale, 24 * scale, seed=3)
    got, hooks = {}, []
    for i, blk in enumerate(net.res_blocks):
        hooks.append(blk.register_forward_hook(lambda m, inp, o, k="res_blocks.%d" % i: got.__setitem__(k, o.detach().float().cpu())))
    with torch.no_grad():
        y = net(lr.cuda()).float().cpu()
    for h in hooks:
        h.remove()
    ref_sd = {k: v.clone() for k, v in sd.items()}
    S.tie_upsampling(ref_sd)
    taps = {}
    with torch.no_grad():
        y_ref = S.generator_forward(ref_sd, lr, scale, n_res, taps)
    tol = TOL[precision]
    worst = max((rel(v, taps[k]), k) for k, v in got.items())
    assert worst[0] < tol, "per-block error %g at %s" % worst
    # end to end (trunk + conv2 / BatchNorm + one or two conv / BatchNorm / shuffle stages + the 9x9 output conv + tanh): the roundings of
    # the bf16 stages behind the last tap add up — bounded by twice the per-layer tolerance, like the full-depth EDSR test
    assert rel(y, y_ref) < 2 * tol
    gsd = net.state_dict()
    for k in sd:                                   # BatchNorm running statistics, incl. the shared up-sampling BatchNorm
        if "running" in k:
            assert rel(gsd[k], ref_sd[k]) < tol, k


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
@pytest.mark.parametrize("case", SRGAN_CASES[:2], ids=lambda c: c[0])
def test_reference_golden_forward(precision, sgolden, case):
    from sradsgan_b200.model.srgan import GeneratorResNet
    name, scale, n_res, batch, lrs = case
    gold = sgolden[name]
    sd = S.tie_upsampling(S.make_state(S.generator_spec(scale, n_res), seed=gold["cfg"]["wseed"], init="fan"))
    net = GeneratorResNet(n_residual_blocks=n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.cuda().train()
    lr, hr = S.synthetic_batch(batch, scale, lrs * scale, seed=gold["cfg"]["dseed"])
    with torch.no_grad():
        y = net(lr.cuda()).float().cpu()
    # batch 2 x (10..12)^2 pixels per channel: BatchNorm divides by a poorly conditioned variance, so bf16 rounding of its input is
    # amplified — the stated tolerance is checked at a well-conditioned batch above; here the reference's own output is the target
    assert rel(y, gold["out"]) < (1e-4 if precision == "fp32" else 3e-2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
def test_trainer_steps_vs_reference_golden(precision, sgolden):
    from sradsgan_b200 import ops
    from sradsgan_b200.model.srgan import SRGAN
    c = sgolden["train_steps"]["cfg"]
    G = S.tie_upsampling(S.make_state(S.generator_spec(c["scale"], c["n_res"]), seed=c["gseed"], init="fan"))
    D = S.make_state(S.discriminator_spec(), seed=c["dseed"], init="fan")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    net = SRGAN(srgan_args(lr=c["lr"], scale_factor=c["scale"], batch_size=c["batch"], vgg_state=V, precision=precision))
    net.n_residual_blocks = c["n_res"]
    net.build(init=False)
    net.generator.load_state_dict(G, strict=True)
    net.discriminator.load_state_dict(D, strict=True)
    ops.bump_weight_generation()
    tol = 1e-3 if precision == "fp32" else 5e-2
    for it, want in enumerate(sgolden["train_steps"]["steps"]):
        lr, hr = S.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        out = net.train_step(lr.cuda(), hr.cuda())
        for k in ("loss_G", "loss_D", "pixel", "content", "adv"):
            assert abs(out[k].item() - want[k]) <= tol * max(1.0, abs(want[k])), (it, k, out[k].item(), want[k])
    if precision == "fp32":
        noise = S.noise_grad_keys(net.generator.state_dict()) | S.noise_grad_keys(net.discriminator.state_dict())
        for name, mod in (("G", net.generator), ("D", net.discriminator)):
            msd = mod.state_dict()
            for k, w in want[name].items():
                if k not in noise and "num_batches" not in k:
                    assert abs(summarize(msd[k].float().cpu(), 8)["norm"] - w["norm"]) <= 2e-3 * max(1e-9, w["norm"]), (name, k)


def test_graphed_step_runs_and_matches_eager_loss():
    """the CUDA-graph replay of the SRGAN iteration (the path train() takes) against the eagerly launched step on the same data"""
    import numpy as np
    from sradsgan_b200 import ops
    from sradsgan_b200.model.srgan import SRGAN
    prev = ops.config.compute_dtype
    try:
        V = O.make_state(O.vgg_spec(), seed=5, init="fan")
        lr, hr = S.synthetic_batch(4, 4, 96, seed=2)
        losses = {}
        for mode in ("eager", "graph"):
            torch.manual_seed(0); np.random.seed(0)
            net = SRGAN(srgan_args(scale_factor=4, batch_size=4, crop_size=96, vgg_state=V, precision="bf16", seed=3))
            net.n_residual_blocks = 3
            net.build(init=True)
            fn = net.train_step if mode == "eager" else net.graphed_step
            for _ in range(3):
                out = fn(lr.cuda(), hr.cuda())
            losses[mode] = (out["loss_G"].item(), out["loss_D"].item())
            assert all(torch.isfinite(torch.tensor(v)) for v in losses[mode])
        assert abs(losses["eager"][0] - losses["graph"][0]) <= 2e-2 * max(1.0, abs(losses["eager"][0]))
        assert abs(losses["eager"][1] - losses["graph"][1]) <= 5e-2 * max(1.0, abs(losses["eager"][1]))
    finally:
        ops.config.compute_dtype = prev
