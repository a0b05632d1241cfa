"""GPU: the discriminator-attention primitives of csrc/cbam.cu through the C ABI against their documented-semantics emulation
(oracle/ops_emu.py, torch CPU fp32), and the composed CBAM (ops.cbam_attention) against the reference's module-by-module
arithmetic (model/base_networks.py:366-457) — forward, gradients and the WGAN-GP style double backward.

Tolerances: fp32 operands 1e-5 relative L2 (summation order); bf16 operands are read exactly, results that are written as bf16
carry one rounding (4e-3); index outputs must match exactly."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(2, 256, 27, 27), (3, 64, 5, 7), (1, 128, 9, 9)]


@pytest.fixture(scope="module")
def be():
    from sradsgan_b200 import _lib
    b = _lib.CudaBackend()
    b.device_check()
    return b


@pytest.fixture(scope="module")
def emu():
    from oracle import ops_emu
    return ops_emu.EmuBackend()


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _full(shape, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_primitives_match_emulation(be, emu, shape, dtype):
    n, c, h, w = shape
    P = h * w
    g = torch.Generator().manual_seed(c + h)
    x, y = _full(shape, dtype, 1), _full(shape, dtype, 2)
    s, s2, a, b = (torch.randn(n, c, generator=g) for _ in range(4))
    m, g0, g1 = (torch.randn(n, P, generator=g) for _ in range(3))
    cidx = torch.randint(0, c, (n, P), generator=g, dtype=torch.int32)
    idx = torch.randint(0, P, (n, c), generator=g, dtype=torch.int32)
    cu = lambda t: None if t is None else t.cuda()
    tol_full = 1e-5 if dtype == torch.float32 else 4e-3
    # elementwise family: each term alone and all together
    for kw in (dict(x=x, s=s, m=m), dict(x=x, m=m), dict(x=x, s=s), dict(s2=s2, g0=g0, g1=g1, cidx=cidx), dict(s2=s2, g1=g1, cidx=cidx),
               dict(a=a, b=b, idx=idx), dict(x=x, s=s, m=m, s2=s2, g0=g0, g1=g1, cidx=cidx, a=a, b=b, idx=idx, acc=y)):
        got = be.cbam_ew(x.cuda(), **{k: cu(v) for k, v in kw.items()})
        ref = emu.cbam_ew(x, **kw)
        assert got.dtype == dtype and rel(got, ref) < tol_full, sorted(kw)
    # reductions (fp32 results)
    for kw in (dict(), dict(b=y), dict(b=y, m=m), dict(m=g0, scale=1.0 / c, g1=g1, cidx=cidx)):
        got = be.cbam_red_c(x.cuda(), **{k: (cu(v) if torch.is_tensor(v) else v) for k, v in kw.items()})
        assert rel(got, emu.cbam_red_c(x, **kw)) < 1e-5, sorted(kw)
    for kw in (dict(), dict(b=y), dict(b=y, s=s), dict(s=s, scale=1.0 / c)):
        got = be.cbam_red_p(x.cuda(), **{k: (cu(v) if torch.is_tensor(v) else v) for k, v in kw.items()})
        assert rel(got, emu.cbam_red_p(x, **kw)) < 1e-5, sorted(kw)
    pooled, pidx = be.cbam_pool_hw(x.cuda())
    pooled_ref, pidx_ref = emu.cbam_pool_hw(x)
    assert rel(pooled, pooled_ref) < 1e-5 and torch.equal(pidx.cpu(), pidx_ref)
    q, qidx = be.cbam_cpool(x.cuda(), s.cuda())
    q_ref, qidx_ref = emu.cbam_cpool(x, s)
    assert rel(q, q_ref) < 1e-5 and torch.equal(qidx.cpu(), qidx_ref)
    assert rel(be.cbam_gather_hw(x.cuda(), idx.cuda()), emu.cbam_gather_hw(x, idx)) < 1e-6
    assert rel(be.cbam_gather_c(x.cuda(), s.cuda(), cidx.cuda()), emu.cbam_gather_c(x, s, cidx)) < 1e-6


def test_pooling_ties_pick_the_first_position(be):
    """bf16 maps hold many exactly equal values: the arg-max must be the first pixel / channel (the adjoint routes the whole
    gradient there, like ATen's max pooling backward)."""
    x = torch.zeros(1, 64, 6, 6).bfloat16()
    x[0, 3, 2, 1] = 1.0
    x[0, 3, 4, 5] = 1.0          # tie over pixels: first = row 2, column 1 -> p = 13
    x[0, 7, 0, 0] = 1.0          # pixel 0: channels 3.. are 0, channel 7 is the unique maximum
    xc = x.contiguous(memory_format=torch.channels_last).cuda()
    pooled, idx = be.cbam_pool_hw(xc)
    assert idx[0, 3].item() == 13 and pooled[1, 0, 3].item() == 1.0 and idx[0, 0].item() == 0
    q, cidx = be.cbam_cpool(xc, torch.ones(1, 64, device="cuda"))
    assert cidx[0, 0].item() == 7 and cidx[0, 13].item() == 3 and cidx[0, 1].item() == 0


def test_small_gemm_strided_views(be):
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(32, 256, generator=g).cuda(), torch.randn(16, 256, generator=g).cuda()
    assert rel(be.small_gemm_nt(a, b), a.cpu() @ b.cpu().t()) < 1e-6
    gy = torch.randn(32, 16, generator=g).cuda()
    assert rel(be.small_gemm_nt(gy, b.t()), gy.cpu() @ b.cpu()) < 1e-6            # b.t(): [256, 16] view with strides (1, 256)
    assert rel(be.small_gemm_nt(gy.t(), a.t()), gy.cpu().t() @ a.cpu()) < 1e-6


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_cbam_attention_matches_module_path_to_second_order(mode):
    """ops.cbam_attention == ChannelAttention -> SpatialAttention evaluated module by module (ATen autograd), including the
    gradient of a gradient-norm penalty with respect to every parameter (the WGAN-GP pattern, model/sradsgan.py:611-639)."""
    from sradsgan_b200 import ops
    from sradsgan_b200.model.sradsgan import ChannelAttention, SpatialAttention
    prev = ops.config.compute_dtype
    ops.set_precision(mode)
    try:
        torch.manual_seed(0)
        ca, sa = ChannelAttention(256).cuda(), SpatialAttention().cuda()
        dt = ops.config.compute_dtype
        x0 = (torch.randn(4, 256, 27, 27, device="cuda") * 0.7).to(dt).contiguous(memory_format=torch.channels_last)
        params = list(ca.parameters()) + list(sa.parameters())

        def run(fused, fp32_truth=False):
            x = (x0.float() if fp32_truth else x0.clone()).requires_grad_(True)
            if fp32_truth:
                ops.set_precision("fp32")
            try:
                y = ops.cbam_attention(x, ca.fc1.weight, ca.fc2.weight, sa.conv1) if fused else sa(ca(x))
            finally:
                ops.set_precision(mode)
            wgt = torch.linspace(-1, 1, y.numel(), device="cuda").view_as(y).to(y.dtype)
            gx, = torch.autograd.grad((y * wgt).sum(), x, create_graph=True)
            pen = ((gx.float().flatten(1).norm(dim=1) - 1) ** 2).mean()
            gp = torch.autograd.grad(pen, params + [x], allow_unused=True)
            return [y.detach(), gx.detach(), pen.detach().reshape(1)] + list(gp)

        fused = run(True)
        eager = run(False)
        truth = run(False, fp32_truth=True) if mode == "bf16" else eager       # fp32 module path on the same (bf16-representable) input
        names = ["y", "dx", "penalty"] + ["d2 %d" % i for i in range(len(params) + 1)]
        for name, a, b, t in zip(names, fused, eager, truth):
            assert (a is None) == (t is None), name
            if a is None:
                continue
            if mode == "fp32":
                assert rel(a, t) < (2e-5 if name in ("y", "dx", "penalty") else 1e-4), name
            else:
                # bf16: within the stated 1e-2 of the fp32 evaluation for first-order quantities; never worse than 1.25x the error
                # of the ATen bf16 path (which rounds s*x and the gates to bf16) for the second-order ones
                e_fused, e_eager = rel(a, t), rel(b, t)
                assert e_fused < max(1e-2 if name in ("y", "penalty") else 1.5e-2, 1.25 * e_eager), (name, e_fused, e_eager)
    finally:
        ops.config.compute_dtype = prev
