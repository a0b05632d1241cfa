"""GPU: the loss-reduction / elementwise-glue kernels (csrc/losses.cu) and the CGAM channel-attention kernels (csrc/cgam.cu)
through the C ABI vs their documented-semantics emulation (oracle/ops_emu.py = the oracle's formulas + torch autograd on CPU).
Reference call sites: model/sradsgan.py:685-688,:834,:838 (L1 / MSE content losses), :46-52 (wgan-gp mean), :611 (GP
interpolates), :623-637 (GP norm + penalty), :178-213 (CGAM)."""
import pytest
import torch

from oracle import ops_emu
from oracle import sradsgan_oracle as O

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


def cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("p", [1, 2])
@pytest.mark.parametrize("case", ["f32_nhwc", "bf16_nhwc", "f32_vs_nchw", "f32_bf16", "ragged"])
def test_diff_mean_forward_backward(be, p, case):
    emu = ops_emu.EmuBackend()
    g = torch.Generator().manual_seed(7)
    shape = (3, 3, 37, 41) if case in ("f32_vs_nchw", "ragged") else (2, 64, 27, 31)
    a = torch.randn(*shape, generator=g)
    b = torch.randn(*shape, generator=g)
    if case == "ragged":
        a, b = a[:, :, :35, :33].contiguous(), b[:, :, :35, :33].contiguous()     # 3*3*35*33 = 10395 elements: not a multiple of 4
    if case == "bf16_nhwc":
        a, b = a.bfloat16(), b.bfloat16()
    if case == "f32_bf16":
        b = b.bfloat16()
    b[0, 0, 0, :5] = a[0, 0, 0, :5].to(b.dtype)                                   # exact zeros of a - b: sign(0) = 0
    ac = cl(a.cuda())
    bc = b.cuda() if case in ("f32_vs_nchw", "ragged") else cl(b.cuda())           # NCHW target (the loader's HR batch) vs NHWC
    got = be.diff_mean(ac, bc, p)
    want = emu.diff_mean(a, b, p)
    assert got.dim() == 0 and abs(got.item() - want.item()) <= 2e-6 * abs(want.item())
    again = be.diff_mean(ac, bc, p)                                               # the workspace ticket re-arms itself; deterministic
    assert again.item() == got.item()
    up = torch.tensor(0.37)
    da = be.diff_mean_bwd(ac, bc, p, up.cuda(), scale=1.5)
    ref = emu.diff_mean_bwd(a, b, p, up, scale=1.5)
    assert da.shape == a.shape and da.dtype == a.dtype
    assert rel(da, ref) < (1e-6 if a.dtype == torch.float32 else 4e-3)
    if a.dtype == b.dtype:
        assert float(da.float().cpu()[0, 0, 0, :5].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mean_forward_backward(be, dtype):
    emu = ops_emu.EmuBackend()
    x = torch.randn(16, 1, 14, 14, generator=torch.Generator().manual_seed(1)).to(dtype)
    got = be.mean(x.cuda(), -1.0)
    assert abs(got.item() - emu.mean(x, -1.0).item()) < 1e-6
    up = torch.tensor(2.5)
    dx = be.mean_bwd(up.cuda(), -1.0, x.cuda())
    assert dx.shape == x.shape and dx.dtype == dtype
    assert rel(dx, emu.mean_bwd(up, -1.0, x)) < (1e-6 if dtype == torch.float32 else 4e-3)


@pytest.mark.parametrize("norm", [0, 1, 2])
@pytest.mark.parametrize("penalty", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gp_penalty_forward_backward(be, norm, penalty, dtype):
    emu = ops_emu.EmuBackend()
    g = torch.Generator().manual_seed(11 + norm)
    grad = (torch.randn(3, 3, 29, 23, generator=g) * 0.9).to(dtype)       # per-pixel norms around 1: both hinge branches
    gc = cl(grad.cuda())
    got = be.gp_penalty(gc, norm, penalty)
    want = emu.gp_penalty(grad, norm, penalty)
    assert abs(got.item() - want.item()) <= 3e-6 * max(abs(want.item()), 1e-3)
    up = torch.tensor(0.8)
    d = be.gp_penalty_bwd(gc, norm, penalty, up.cuda(), scale=11.0)
    ref = emu.gp_penalty_bwd(grad, norm, penalty, up, scale=11.0)
    assert d.dtype == dtype and rel(d, ref) < (2e-6 if dtype == torch.float32 else 5e-3)


@pytest.mark.parametrize("real_nchw", [True, False])
def test_lerp_and_layout_kernels(be, real_nchw):
    emu = ops_emu.EmuBackend()
    g = torch.Generator().manual_seed(5)
    real = torch.rand(4, 3, 19, 22, generator=g)
    fake = torch.randn(4, 3, 19, 22, generator=g)
    alpha = torch.rand(4, 1, 1, 1, generator=g)
    rc = real.cuda() if real_nchw else cl(real.cuda())
    out = be.lerp(rc, cl(fake.cuda()), alpha.cuda(), torch.float32)
    assert out.is_contiguous(memory_format=torch.channels_last)
    assert rel(out, emu.lerp(real, fake, alpha, torch.float32)) < 2e-7                  # the same fp32 expression (up to FMA contraction)
    y = be.nchw_to_nhwc(real.cuda(), torch.float32)
    assert y.is_contiguous(memory_format=torch.channels_last) and torch.equal(y.cpu(), real)
    y16 = be.nchw_to_nhwc(real.cuda(), torch.bfloat16)
    assert torch.equal(y16.cpu(), real.bfloat16())
    a = torch.randn(2, 64, 9, 7, generator=g)
    b = torch.randn(2, 64, 9, 7, generator=g).bfloat16()
    s = be.add_cast(cl(a.cuda()), cl(b.cuda()), torch.bfloat16)
    assert torch.equal(s.cpu(), (a + b.float()).bfloat16())
    c = be.add_cast(cl(a.cuda()), None, torch.bfloat16)
    assert torch.equal(c.cpu(), a.bfloat16())
    odd = torch.randn(1, 3, 5, 7, generator=g)                                       # 105 elements: scalar tail
    assert torch.equal(be.add_cast(cl(odd.cuda()), None, torch.bfloat16).cpu(), odd.bfloat16())


@pytest.mark.parametrize("shape", [(2, 54, 54), (3, 13, 17), (1, 128, 128), (2, 5, 3)])
@pytest.mark.parametrize("scale", [0.05, 1.0])
def test_cgam_forward_backward(be, shape, scale):
    """`scale` sets the magnitude of the activations: 0.05 gives a soft attention (logits O(1)), 1.0 a nearly one-hot one
    (gram entries ~ +-H*W)."""
    n, h, w = shape
    g = torch.Generator().manual_seed(h * w + 1)
    x = torch.randn(n, 64, h, w, generator=g) * scale
    gamma = torch.tensor([0.5])
    dy = torch.randn(n, 64, h, w, generator=g)
    # ground truth in float64 (two fp32 evaluations of a near one-hot softmax over logits of magnitude ~H*W differ by ~1e-4)
    xs = x.double().requires_grad_(True)
    gs = gamma.double().requires_grad_(True)
    y_ref, A_ref = ops_emu.EmuBackend._cgam_math(xs, gs)
    dx_ref, dg_ref = torch.autograd.grad(y_ref, [xs, gs], dy.double())
    soft = scale < 0.5
    xc = cl(x.cuda())
    y32, y16, A = be.cgam_fwd(xc, gamma.cuda(), torch.bfloat16)
    assert rel(A, A_ref) < (2e-5 if soft else 1e-3)
    assert rel(y32, y_ref) < (2e-6 if soft else 2e-4)
    assert torch.equal(y16.cpu(), y32.cpu().bfloat16())
    with torch.no_grad():
        e_oracle = rel(O.cgam({"c.gamma": gamma}, "c", x), y_ref)                   # the oracle's own fp32 evaluation vs the truth
    print("cgam %s scale %g: kernel %.2e, oracle fp32 %.2e (vs float64)" % (shape, scale, rel(y32, y_ref), e_oracle))
    dx, dg = be.cgam_bwd(cl(dy.cuda()), xc, A, gamma.cuda())
    assert rel(dx, dx_ref) < (1e-5 if soft else 2e-3)
    budget = (dy.abs() * ((y_ref.detach() - x.double()) / 0.5).abs()).sum().item()  # sum of |terms| of dgamma = <dy, A X>
    assert abs(dg.item() - dg_ref.item()) <= 2e-5 * budget
    acc = torch.full((1,), 3.0).cuda()
    dx2, none = be.cgam_bwd(cl(dy.cuda()), xc, A, gamma.cuda(), dgamma_into=acc)
    assert none is None and abs(acc.item() - (3.0 + dg.item())) <= 1e-5 * max(1.0, abs(dg.item()))
    assert torch.equal(dx2, dx)
