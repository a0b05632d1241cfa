"""GPU: the SRAGAN sibling on the CUDA path — BasicBlock = the conv-pair node on the tcgen05 halo kernel (pooling
partials from its epilogue) + the fused local-attention chain, CAM / PAM = the CGAM / SGAM kernels, BatchNorm on the fused kernels —
against the CPU oracle (oracle/sragan_oracle.py) and the golden vectors recorded from the UNMODIFIED reference `model.sragan`.

Tolerances (BASELINE.json north_star): per-layer relative L2 error <= 1e-4 in fp32 mode, <= 1e-2 in bf16 mode."""
import os

import numpy as np
import pytest
import torch

from oracle import sradsgan_oracle as O
from oracle import sragan_oracle as A
from oracle.make_golden import summarize
from test_sragan_cpu import sragan_args

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = {"fp32": 1e-4, "bf16": 1e-2}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def agolden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "sragan_golden.pt"), weights_only=False)


@pytest.fixture()
def precision(request):
    from sradsgan_b200 import ops
    prev = ops.config.compute_dtype
    ops.set_precision(request.param)
    yield request.param
    ops.config.compute_dtype = prev


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
@pytest.mark.parametrize("scale,n_basic", [(4, 3), (3, 2)])
def test_generator_per_block_parity(precision, scale, n_basic):
    from sradsgan_b200.model.sragan import GeneratorResNet, ResidualBlock_Block_WithAttention
    n_res = 3
    sd = A.tie_upsampling(A.make_state(A.generator_spec(scale, n_res, n_basic), seed=13 + scale, init="fan"))
    net = GeneratorResNet(ResidualBlock_Block_WithAttention, n_residual_blocks=n_res, n_basic_blocks=n_basic, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.cuda().train()
    lr, hr = A.synthetic_batch(8, scale, 24 * scale, seed=3)
    got, hooks = {}, []
    for i, blk in enumerate(net.res_blocks):
        hooks.append(blk.register_forward_hook(lambda m, inp, o, k="res_blocks.%d" % i: got.__setitem__(k, o.detach().float().cpu())))
    with torch.no_grad():
        y = net(lr.cuda()).float().cpu()
    for h in hooks:
        h.remove()
    ref_sd = {k: v.clone() for k, v in sd.items()}
    A.tie_upsampling(ref_sd)
    taps = {}
    with torch.no_grad():
        y_ref = A.generator_forward(ref_sd, lr, scale, n_res, n_basic, taps)
    tol = TOL[precision]
    worst = max((rel(v, taps[k]), k) for k, v in got.items())
    assert worst[0] < tol, "per-block error %g at %s" % worst
    # Behind the last tap: conv2 / BN, CAM, PAM, 1x1, one or two conv / BN / shuffle stages, conv3, tanh.  With these O(1) synthetic
    # activations CAM's softmax(rowmax(E) - E) over a 576-pixel gram is nearly one-hot, i.e. badly conditioned: it amplifies the trunk's
    # (in-tolerance) bf16 error.  fp32 mode is held to 2x the per-layer tolerance; bf16 mode to the north-star's END-TO-END criterion, the
    # PSNR against the HR batch within 0.01 dB, and to a relative error that would expose a wrong tail (a sign / layout / stage error is O(1)).
    if precision == "fp32":
        assert rel(y, y_ref) < 2 * tol
    else:
        assert abs(O.psnr(y, hr) - O.psnr(y_ref, hr)) <= 0.01
        assert rel(y, y_ref) < 8e-2


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
def test_residual_block_backward_parity(precision):
    """gradients of ONE ResidualBlock_Block_WithAttention (three BasicBlocks + its own attention tail) — the code that is new in this
    sibling: the conv-pair node on 64 -> 64 -> 64, the fused local-attention chain behind it, the activation after the residual add."""
    from sradsgan_b200.model.sragan import BasicBlock, ResidualBlock_Block_WithAttention
    from collections import OrderedDict
    spec = OrderedDict()
    for j in range(2):
        A._block(spec, "blocks.%d" % j)
    A._block(spec, "last_conv")
    A._la(spec, "")
    spec = OrderedDict((k.lstrip("."), v) for k, v in spec.items())
    sd = A.make_state(spec, seed=17, init="fan")
    blk = ResidualBlock_Block_WithAttention(BasicBlock, n_blocks=3, nc=64, norm_type=None, act_type='lrelu')
    assert list(blk.state_dict().keys()) == list(sd.keys())
    blk.load_state_dict(sd, strict=True)
    blk.cuda().train()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(8, 64, 24, 24, generator=g) * 0.5
    xg = x.clone().cuda().requires_grad_(True)
    y = blk(xg)
    (y.float() ** 2).mean().backward()
    mine = {("res." + k): v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    y_ref = A.res_block(mine, "res", xr, 3)
    (y_ref ** 2).mean().backward()
    tol = TOL[precision]
    assert rel(y, y_ref) < tol
    # bf16: the same bound as the CLAM / SLAM gradients of the SRADSGAN blocks (tests/test_gpu_model_parity.py) — the max-pooling routes
    # of four attention chains are discrete (an arg-max that flips under bf16 rounding moves a whole gradient contribution), on top of the
    # bf16 roundings of six input-gradient convolutions; fp32 mode holds the wiring itself to 2e-3
    gtol = 2e-3 if precision == "fp32" else 1.5e-1
    assert rel(xg.grad, xr.grad) < gtol
    for k, p in blk.named_parameters():
        if p.dim() == 4 and p.shape[-1] == 3:                    # the 3x3 convolution weights carry the bulk of the gradient
            assert rel(p.grad, mine["res." + k].grad) < gtol, k


@pytest.mark.parametrize("precision", ["fp32", "bf16"], indirect=True)
def test_trainer_steps_vs_reference_golden(precision, agolden):
    from sradsgan_b200 import ops
    from sradsgan_b200.model.sragan import SRAGAN
    c = agolden["train_steps"]["cfg"]
    G = A.tie_upsampling(A.make_state(A.generator_spec(c["scale"], c["n_res"], c["n_basic"]), seed=c["gseed"], init="fan"))
    D = O.make_state(O.discriminator_spec(), seed=c["dseed"], init="ref")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    net = SRAGAN(sragan_args(scale_factor=c["scale"], batch_size=c["batch"], vgg_state=V, precision=precision))
    net.n_residual_blocks, net.n_basic_blocks = c["n_res"], c["n_basic"]
    net.build(init=False)
    net.generator.load_state_dict(G, strict=True)
    net.discriminator.load_state_dict(D, strict=True)
    ops.bump_weight_generation()
    tol = 1e-3 if precision == "fp32" else 5e-2
    for it, want in enumerate(agolden["train_steps"]["steps"]):
        lr, hr = A.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        np.random.seed(c["np_seed"] + it)
        net._alpha_override = torch.Tensor(np.random.random((c["batch"], 1, 1, 1))).cuda()
        out = net.train_step(lr.cuda(), hr.cuda())
        for k in ("loss_G", "pixel", "content", "gp"):
            assert abs(out[k].item() - want[k]) <= tol * max(1.0, abs(want[k])), (it, k, out[k].item(), want[k])


def test_graphed_step_runs():
    """the CUDA-graph replay of the SRAGAN iteration (the path train() takes): finite losses, same ballpark as the eager step"""
    from sradsgan_b200 import ops
    from sradsgan_b200.model.sragan import SRAGAN
    prev = ops.config.compute_dtype
    try:
        V = O.make_state(O.vgg_spec(), seed=5, init="fan")
        lr, hr = A.synthetic_batch(4, 4, 96, seed=2)
        res = {}
        for mode in ("eager", "graph"):
            torch.manual_seed(0); np.random.seed(0)
            net = SRAGAN(sragan_args(scale_factor=4, batch_size=4, crop_size=96, vgg_state=V, precision="bf16", seed=3))
            net.n_residual_blocks, net.n_basic_blocks = 2, 2
            net.build(init=True)
            fn = net.train_step if mode == "eager" else net.graphed_step
            for _ in range(2):
                out = fn(lr.cuda(), hr.cuda())
            res[mode] = (out["loss_G"].item(), out["pixel"].item())
            assert all(np.isfinite(v) for v in res[mode])
        assert abs(res["eager"][1] - res["graph"][1]) <= 2e-2 * max(1.0, abs(res["eager"][1]))
    finally:
        ops.config.compute_dtype = prev
