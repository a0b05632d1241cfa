"""GPU: the EDSR workload (BASELINE.json configs[4]) on the CUDA path — modules -> autograd Functions ->
C ABI -> the same sm_100a convolution kernels as SRADSGAN — against the CPU oracle (oracle/edsr_oracle.py) and the golden
vectors recorded from the UNMODIFIED reference `model.edsr.Net`.

Tolerances (BASELINE.json north_star): per-layer relative L2 error <= 1e-4 in fp32 mode, <= 1e-2 in bf16 mode."""
import os

import pytest
import torch

from oracle import edsr_oracle as E
from oracle.make_golden import summarize
from oracle.make_golden_edsr import EDSR_CASES
from test_edsr_cpu import edsr_args

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def egolden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "edsr_golden.pt"), weights_only=False)


@pytest.fixture()
def precision(request):
    from sradsgan_b200 import ops
    prev = ops.config.compute_dtype
    ops.set_precision(request.param)
    yield request.param
    ops.config.compute_dtype = prev


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
@pytest.mark.parametrize("case", EDSR_CASES, ids=lambda c: c[0])
def test_per_layer_parity_and_golden(precision, egolden, case):
    from sradsgan_b200.model.edsr import Net
    name, scale, n_res, batch, lrs = case
    gold = egolden[name]
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(scale, n_res), seed=gold["cfg"]["wseed"], init="fan"))
    net = Net(3, 256, n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.cuda()
    lr, hr = E.synthetic_batch(batch, scale, lrs * scale, seed=gold["cfg"]["dseed"])
    got, hooks = {}, []
    for i, blk in enumerate(net.residual_layers):
        hooks.append(blk.register_forward_hook(lambda m, inp, o, k="residual_layers.%d" % i: got.__setitem__(k, o.detach().float().cpu())))
    with torch.no_grad():
        y = net(lr.cuda()).float().cpu()
    for h in hooks:
        h.remove()
    taps = {}
    with torch.no_grad():
        y_ref = E.edsr_forward(sd, lr, scale, n_res, taps)
    tol = TOL[precision]
    worst = max((rel(v, taps[k]), k) for k, v in got.items())
    assert worst[0] < tol, "per-layer error %g at %s" % worst
    assert rel(y, y_ref) < tol
    assert rel(y, gold["out"]) < tol                                  # the reference's own output
    loss = (y - hr).abs().mean().item()
    assert abs(loss - gold["loss"]) <= (1e-4 if precision == "fp32" else 1e-2) * max(1.0, abs(gold["loss"]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
def test_backward_parity(precision):
    from sradsgan_b200.model.edsr import Net
    scale, n_res = 4, 2
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(scale, n_res), seed=5, init="fan"))
    net = Net(3, 256, n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.cuda()
    lr, hr = E.synthetic_batch(2, scale, 64, seed=9)
    y = net(lr.cuda())
    # smooth loss: an L1 loss's sign() gradient flips with the output's rounding and makes the check ill-posed
    (0.5 * (y.float() - hr.cuda()) ** 2).mean().backward()
    mine = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    E.tie_upsampling(mine)
    (0.5 * (E.edsr_forward(mine, lr, scale, n_res) - hr) ** 2).mean().backward()
    tol = 1e-3 if precision == "fp32" else 5e-2
    worst = max((rel(p.grad, mine[k].grad), k) for k, p in net.named_parameters())
    assert worst[0] < tol, "gradient error %g at %s" % worst


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_training_steps_vs_reference_golden(egolden, prec):
    """two full iterations (forward, L1, backward through the fused ResnetBlock nodes, flat fused Adam) == the reference's
    recorded losses / parameters"""
    from sradsgan_b200 import ops
    from sradsgan_b200.model.edsr import EDSR
    c = egolden["train_steps"]["cfg"]
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(c["scale"], c["n_res"]), seed=c["wseed"], init="fan"))
    prev = ops.config.compute_dtype
    try:
        net = EDSR(edsr_args(lr=c["lr"], scale_factor=c["scale"], batch_size=c["batch"], precision=prec))
        net.num_residuals = c["n_res"]
        net.build(init=False)
        net.generator.load_state_dict(sd, strict=True)
        ops.bump_weight_generation()
        ltol = 5e-4 if prec == "fp32" else 3e-2
        for it, want in enumerate(egolden["train_steps"]["steps"]):
            lr, hr = E.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
            out = net.train_step(lr.cuda(), hr.cuda())
            assert abs(out["loss_G"].item() - want["loss_G"]) <= ltol * max(1.0, abs(want["loss_G"])), (it, out["loss_G"].item(), want["loss_G"])
            if prec == "fp32":
                gsd = net.generator.state_dict()
                for k, w in want["params"].items():
                    assert abs(summarize(gsd[k].cpu(), 8)["norm"] - w["norm"]) <= 5e-4 * max(1e-6, w["norm"]), (it, k)
    finally:
        ops.config.compute_dtype = prev


def test_graphed_step_matches_eager_steps(egolden):
    from sradsgan_b200 import ops
    from sradsgan_b200.model.edsr import EDSR
    c = egolden["train_steps"]["cfg"]
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(c["scale"], c["n_res"]), seed=c["wseed"], init="fan"))
    prev = ops.config.compute_dtype
    try:
        nets = []
        for _ in range(2):
            net = EDSR(edsr_args(lr=c["lr"], scale_factor=c["scale"], batch_size=c["batch"], precision="bf16"))
            net.num_residuals = c["n_res"]
            net.build(init=False)
            net.generator.load_state_dict(sd, strict=True)
            ops.bump_weight_generation()
            nets.append(net)
        eager, graphed = nets
        for it in range(3):
            lr, hr = E.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
            a = eager.train_step(lr.cuda(), hr.cuda())["loss_G"].item()
            b = graphed.graphed_step(lr.cuda(), hr.cuda())["loss_G"].item()
            assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (it, a, b)
        assert graphed._graph["launches"] > 0
        assert rel(graphed.optimizer_G.flat_param, eager.optimizer_G.flat_param) < 2e-3
        assert int(graphed.optimizer_G.step_t.item()) == 3 and graphed.optimizer_G.step_count == 3
    finally:
        ops.config.compute_dtype = prev


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
def test_full_depth_forward_parity(precision):
    """the real depth — EDSR(256 filters, 32 residual blocks), x4 — on a small map.  PER-LAYER error (every block fed the
    oracle's own input): the north-star tolerance.  END-TO-END error through all 32 blocks (reference init: each branch is as
    large as its trunk, so rounding noise adds up like sqrt(depth)): 2x that (measured 1.03e-2 at block 30 in bf16)."""
    from sradsgan_b200.model.edsr import Net
    scale, n_res = 4, 32
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(scale, n_res), seed=11, init="ref"))
    net = Net(3, 256, n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.cuda()
    assert sum(p.numel() for p in net.parameters()) == sum(p.numel() for p in E.unique_params(sd))
    lr, _ = E.synthetic_batch(2, scale, 48, seed=13)
    got, hooks = {}, []
    for i, blk in enumerate(net.residual_layers):
        hooks.append(blk.register_forward_hook(lambda m, inp, o, k="residual_layers.%d" % i: got.__setitem__(k, o.detach().float().cpu())))
    with torch.no_grad():
        y = net(lr.cuda()).float().cpu()
    for h in hooks:
        h.remove()
    taps = {}
    with torch.no_grad():
        y_ref = E.edsr_forward(sd, lr, scale, n_res, taps)
    tol = TOL[precision]
    worst = max((rel(v, taps[k]), k) for k, v in got.items())
    assert worst[0] < 2 * tol, "end-to-end error %g at %s" % worst
    assert rel(y, y_ref) < 2 * tol
    with torch.no_grad():
        for i in (1, 8, 16, 24, 31):                     # block i alone, on the oracle's input of that block
            x_in = taps["residual_layers.%d" % (i - 1)].cuda().contiguous(memory_format=torch.channels_last)
            out = net.residual_layers[i](x_in).float().cpu()
            e = rel(out, taps["residual_layers.%d" % i])
            assert e < tol, "per-layer error %g at block %d" % (e, i)
