"""GPU: the fused non-convolution kernels through the C ABI vs their documented-semantics emulation
(oracle/ops_emu.py, which is the oracle's CLAM/SLAM/conv math + torch autograd on CPU)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ops_emu
from sradsgan_b200._lib import ACT_LRELU, ACT_RELU, conv_geom

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from sradsgan_b200 import _lib
    b = _lib.CudaBackend()
    b.device_check()
    return b


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _chain_inputs(n, h, w, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 64, h, w, generator=g).to(dtype)
    t = torch.randn(n, 64, h, w, generator=g)
    fc1 = torch.randn(4, 64, 1, 1, generator=g) * 0.3
    fc2 = torch.randn(64, 4, 1, 1, generator=g) * 0.3
    w7 = torch.randn(1, 2, 7, 7, generator=g) * 0.2
    W = torch.randn(64, 64, 1, 1, generator=g) * 0.125
    b = torch.randn(64, generator=g) * 0.1
    return x, t, fc1, fc2, w7, W, b


@pytest.mark.parametrize("shape", [(2, 54, 54), (1, 13, 17), (3, 24, 24), (1, 7, 5)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_la_chain_forward_backward(be, shape, dtype):
    n, h, w = shape
    x, t, fc1, fc2, w7, W, b = _chain_inputs(n, h, w, dtype, seed=h * w)
    emu = ops_emu.EmuBackend()
    z_ref, _, sv_ref = emu.la_chain_fwd(x, t, fc1, fc2, w7, W, b, want_lowp=False)
    cu = [v.cuda() for v in (x, t, fc1, fc2, w7, W, b)]
    cu[0] = cu[0].contiguous(memory_format=torch.channels_last)
    z32, z16, sv = be.la_chain_fwd(*cu, want_lowp=True)
    # fp32 mode: all chain math in fp32.  bf16 mode: the 64x64 products run on the tensor cores with every fp32 operand split
    # into a hi + lo bf16 pair (three MMAs per product, ~2^-16 relative), fp32 accumulation.
    tol = 2e-5 if dtype == torch.float32 else 1e-4
    assert rel(z32, z_ref) < tol
    assert rel(z16, z_ref) < (1e-6 if dtype == torch.float32 else 4e-3)
    g = torch.Generator().manual_seed(3)
    gz32 = torch.randn(n, 64, h, w, generator=g)
    gz16 = torch.randn(n, 64, h, w, generator=g).to(dtype)
    ref = emu.la_chain_bwd(gz32, gz16, x, sv_ref, fc1, fc2, w7, W)
    got = be.la_chain_bwd(gz32.cuda(), gz16.cuda(), cu[0], sv, cu[2], cu[3], cu[4], cu[5])
    names = ["dx", "d_fc1", "d_fc2", "d_w7", "dW", "db", "dz"]
    for nm, a, r in zip(names, got, ref):
        lim = 2e-4 if (dtype == torch.float32 or nm == "dz") else (4e-3 if nm == "dx" else 1e-3)
        assert rel(a, r) < lim, (nm, rel(a, r))
    # only the fp32 gradient present: dz is the incoming gradient itself
    got2 = be.la_chain_bwd(gz32.cuda(), None, cu[0], sv, cu[2], cu[3], cu[4], cu[5])
    ref2 = emu.la_chain_bwd(gz32, None, x, sv_ref, fc1, fc2, w7, W)
    assert rel(got2[4], ref2[4]) < (1e-3 if dtype == torch.bfloat16 else 2e-4)
    assert rel(got2[0], ref2[0]) < (4e-3 if dtype == torch.bfloat16 else 2e-4)


@pytest.mark.parametrize("r,act", [(1, ACT_LRELU), (2, ACT_LRELU), (3, ACT_LRELU), (1, ACT_RELU)])
def test_act_bwd(be, r, act):
    emu = ops_emu.EmuBackend()
    g = conv_geom((2, 64, 10, 12), (64 * r * r, 64, 3, 3), 1, 1)
    gen = torch.Generator().manual_seed(r)
    y = torch.randn(2, 64, 10 * r, 12 * r, generator=gen).bfloat16()
    gy = torch.randn(2, 64, 10 * r, 12 * r, generator=gen)
    ref = emu.act_bwd(gy, y, act, 0.2, r, g, torch.bfloat16)
    got = be.act_bwd(gy.cuda(), y.cuda(), act, 0.2, r, g, torch.bfloat16)
    assert got.shape == ref.shape and rel(got, ref) < 1e-6


@pytest.mark.parametrize("shape", [(4, 64, 20, 24), (2, 128, 27, 27), (16, 512, 14, 14), (3, 256, 7, 9)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_leaky_relu_forward_backward(be, shape, dtype):
    """train-mode BatchNorm2d + LeakyReLU(0.2): batch statistics, running-stat update, first-order backward"""
    emu = ops_emu.EmuBackend()
    g = torch.Generator().manual_seed(shape[1])
    n, c, h, w = shape
    x = (torch.randn(shape, generator=g) * 0.7 + 0.3).to(dtype)
    gamma = 1 + 0.1 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    rm, rv = torch.zeros(c), torch.ones(c)
    y_ref, save_ref = emu.bn_act_fwd(x, gamma, beta, rm, rv, 1e-5, 0.1, 0.2)
    rm_c, rv_c = torch.zeros(c).cuda(), torch.ones(c).cuda()
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    y, save = be.bn_act_fwd(xc, gamma.cuda(), beta.cuda(), rm_c, rv_c, 1e-5, 0.1, 0.2)
    tol = 1e-5 if dtype == torch.float32 else 4e-3
    assert rel(y, y_ref) < tol
    assert rel(save, save_ref) < 1e-4
    assert rel(rm_c, rm) < 1e-4 and rel(rv_c, rv) < 1e-4
    gy = torch.randn(shape, generator=g).to(dtype)
    dx_ref, dg_ref, db_ref = emu.bn_act_bwd(gy, x, save_ref, 0.2)
    dx, dg, db = be.bn_act_bwd(gy.cuda(), xc, save, 0.2)
    # the batch statistics are reduced with fp32 atomics (run-to-run summation order), so an element whose
    # pre-activation is within rounding of 0 may land on the other LeakyReLU branch: leave those out
    z_ref = x.float() * save_ref[2].view(1, -1, 1, 1) + save_ref[3].view(1, -1, 1, 1)
    keep = (z_ref.abs() > 1e-4).float()
    assert rel(dx.cpu().float() * keep, dx_ref.float() * keep) < (3e-4 if dtype == torch.float32 else 6e-3)
    assert rel(dg, dg_ref) < 5e-4 and rel(db, db_ref) < 5e-4
    # double backward (WGAN-GP): cotangent u of dx -> d_gy, d_x, d_gamma; checker = autograd through a
    # differentiable restatement of the first-order backward (oracle/ops_emu.py), kernel = closed form
    u = torch.randn(shape, generator=g).to(dtype)
    r_gy, r_x, r_gamma = emu.bn_act_bwd_bwd(u, gy, x, save_ref, dg_ref, db_ref, 0.2)
    d_gy, d_x, d_gamma = be.bn_act_bwd_bwd(u.cuda(), gy.cuda(), xc, save, dg, db, 0.2)
    lim = 3e-4 if dtype == torch.float32 else 8e-3
    e_gy, e_x = rel(d_gy.cpu().float() * keep, r_gy.float() * keep), rel(d_x.cpu().float() * keep, r_x.float() * keep)
    assert e_gy < lim and e_x < lim, (e_gy, e_x)
    assert rel(d_gamma, r_gamma) < 1e-3, rel(d_gamma, r_gamma)


def test_discriminator_fused_double_backward_matches_unfused():
    """The WGAN-GP penalty and its parameter gradients through the fused any-order Functions (conv+LReLU epilogue,
    BatchNorm+LReLU kernels, closed-form BatchNorm double backward) == the module-by-module differentiable path."""
    from oracle import sradsgan_oracle as O
    from sradsgan_b200 import ops
    from sradsgan_b200.model.sradsgan import Discriminator
    prev = ops.config.compute_dtype
    ops.set_precision("fp32")
    try:
        sd = O.make_state(O.discriminator_spec(), seed=3, init="fan")
        x = torch.rand(2, 3, 48, 48, generator=torch.Generator().manual_seed(1)).cuda()
        res = []
        for unfused in (False, True):
            D = Discriminator()
            D.load_state_dict(sd, strict=True)
            D.cuda()
            ops.config.double_backward = unfused
            xi = x.clone().requires_grad_(True)
            d = D(xi)
            gr = torch.autograd.grad(d, xi, torch.ones_like(d), create_graph=True, retain_graph=True)[0]
            gp = ((gr.float().norm(2, 1) - 1) ** 2).mean()
            gp.backward()
            res.append((gp.item(), {k: (p.grad.clone() if p.grad is not None else None) for k, p in D.named_parameters()},
                        {k: v.clone() for k, v in D.state_dict().items() if "running" in k}))
        ops.config.double_backward = False
        (gp_f, gr_f, st_f), (gp_u, gr_u, st_u) = res
        assert abs(gp_f - gp_u) < 1e-5 * max(1.0, abs(gp_u))
        for k, gu in gr_u.items():
            if k in O.NOISE_GRAD_KEYS:
                continue
            if gr_f[k] is None or gu is None:
                assert (gu is None or gu.abs().max() == 0) and (gr_f[k] is None or gr_f[k].abs().max() == 0), k
                continue
            assert rel(gr_f[k], gu) < 2e-3, (k, rel(gr_f[k], gu))
        for k in st_u:
            assert rel(st_f[k], st_u[k]) < 1e-4, k
    finally:
        ops.config.double_backward = False
        ops.config.compute_dtype = prev


@pytest.mark.parametrize("shape", [(2, 54, 54), (1, 13, 11), (3, 24, 24), (1, 40, 33)])
def test_sgam_flash_attention_forward_backward(be, shape):
    """csrc/sgam.cu (statistics pass, tcgen05 weight x value kernel, tcgen05 + SIMT dS kernel) vs the explicit
    softmax(Q^T K) formulation of the reference (model/sradsgan.py:164-176) in fp32 torch, forward and all gradients."""
    from sradsgan_b200 import ops
    n, h, w = shape
    g = torch.Generator().manual_seed(h * w)
    q = (torch.randn(n, 8, h, w, generator=g) * 1.2).bfloat16()
    k = (torch.randn(n, 8, h, w, generator=g) * 1.2).bfloat16()
    v = torch.randn(n, 64, h, w, generator=g).bfloat16()
    x = torch.randn(n, 64, h, w, generator=g)
    gamma = torch.tensor([0.7])
    dy = torch.randn(n, 64, h, w, generator=g)

    def ref(q, k, v, x, gamma):
        qf = q.flatten(2).permute(0, 2, 1)
        att = torch.softmax(torch.bmm(qf, k.flatten(2)), dim=-1)
        out = torch.bmm(v.flatten(2), att.permute(0, 2, 1)).view(n, 64, h, w)
        return gamma * out + x, out

    leaves = [t.float().clone().requires_grad_(True) for t in (q, k, v, x, gamma)]
    y_ref, out_ref = ref(*leaves)
    y_ref.backward(dy)
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
    cu = [cl(q).requires_grad_(True), cl(k).requires_grad_(True), cl(v).requires_grad_(True), cl(x).requires_grad_(True),
          gamma.cuda().requires_grad_(True)]
    y = ops.SGAMAttention.apply(*cu)
    y.backward(cl(dy))
    torch.cuda.synchronize()
    assert rel(y - cu[3], (y_ref - leaves[3])) < 6e-3          # the attention term itself (bf16 weights and values)
    assert rel(y, y_ref) < 3e-3
    names = ["dq", "dk", "dv", "dx", "dgamma"]
    tols = [2e-2, 2e-2, 8e-3, 1e-6, 3e-2]      # dgamma = sum dy*o is a cancelling sum of bf16-rounded o: looser
    for nm, a, b, tol in zip(names[:4], cu, leaves, tols):
        assert rel(a.grad.float(), b.grad) < tol, (nm, rel(a.grad.float(), b.grad))
    # dgamma = sum dy * o is a cancelling sum over bf16-rounded o: judge it against the sum of magnitudes
    mag = (dy.abs() * out_ref.detach().abs()).sum().item()
    assert abs(cu[4].grad.item() - leaves[4].grad.item()) < 2e-3 * mag


@pytest.mark.parametrize("shape", [(2, 64, 16, 20), (1, 128, 7, 9), (3, 8, 54, 54)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_maxpool2x2_forward_backward(be, shape, dtype):
    """sr_maxpool2x2_fwd / _bwd == torch's max_pool2d and its autograd, bit for bit — including windows with tied maxima
    (values quantised to a few levels: the gradient must go to the FIRST maximum in row-major order) and odd sizes
    (the last row / column belongs to no window: zero gradient)."""
    g = torch.Generator().manual_seed(shape[1] + shape[2])
    x = (torch.randint(0, 5, shape, generator=g).float() * 0.25).to(dtype)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    y = be.maxpool2x2_fwd(xc)
    xr = x.float().requires_grad_(True)
    y_ref = F.max_pool2d(xr, 2, 2)
    assert y.shape == y_ref.shape and torch.equal(y.float().cpu(), y_ref.detach())
    gy = torch.randn(y_ref.shape, generator=g).to(dtype)
    (dx_ref,) = torch.autograd.grad(y_ref, xr, gy.float())
    dx = be.maxpool2x2_bwd(gy.cuda(), xc)
    assert dx.shape == x.shape and torch.equal(dx.float().cpu(), dx_ref)
