"""GPU: per-kernel parity of the convolution family through the C ABI (sradsgan_b200._lib.CudaBackend ->
libsradsgan_b200.so) against torch CPU fp32 on the same (bf16-rounded) operands.

Tolerances: fp32 SIMT path 1e-5 relative L2 (fp32 accumulate, different summation order);
bf16 paths 4e-3 (one bf16 rounding of the output, 2^-9 = 2e-3, fp32 accumulation in both)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from sradsgan_b200 import _lib
    b = _lib.CudaBackend()
    b.device_check()
    return b


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _mk(n, cin, h, w, cout, k, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, cin, h, w, generator=g).to(dtype)
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    return x, wt, b


def _ref_w(wt, dtype):
    return wt.to(dtype).float()


from sradsgan_b200._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, IMPL_AUTO, IMPL_HALO, IMPL_SIMT, IMPL_TCGEN05, conv_geom)

SIMT_CASES = [
    # n, cin, h, w, cout, k, stride, pad
    (2, 3, 20, 20, 64, 3, 1, 1),       # conv1 / MSB.conv1 / D.model.0 / VGG.0 (K5)
    (2, 64, 18, 18, 3, 3, 1, 1),       # conv3 (K6)
    (2, 2, 17, 19, 1, 7, 1, 3),        # SLAM 7x7 (K8)
    (2, 3, 11, 11, 64, 1, 1, 0),       # MSB 1x1
    (1, 64, 16, 16, 64, 3, 2, 1),      # D stride-2
    (1, 64, 15, 15, 8, 1, 1, 0),       # SGAM q/k
    (1, 512, 6, 6, 1, 3, 1, 1),        # D output conv
    (1, 192, 9, 9, 64, 1, 1, 0),       # MSB.conv
]


@pytest.mark.parametrize("case", SIMT_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_simt_fwd_dgrad_wgrad(be, case, dtype):
    n, cin, h, w, cout, k, s, p = case
    x, wt, b = _mk(n, cin, h, w, cout, k, dtype, seed=cin + cout)
    g = conv_geom(x.shape, wt.shape, s, p)
    tol = 1e-5 if dtype == torch.float32 else 4e-3
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    wp = be.pack_weights(wt.cuda(), 0, dtype)
    y = be.conv_fwd(xc, wp, b.cuda(), None, g, ACT_LRELU, 0.2, impl=IMPL_SIMT)
    y_ref = F.leaky_relu(F.conv2d(x.float(), _ref_w(wt, dtype), b, stride=s, padding=p), 0.2)
    assert y.shape == y_ref.shape and rel(y, y_ref) < tol
    gy = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(1)).to(dtype)
    gyc = gy.cuda().contiguous(memory_format=torch.channels_last)
    dx = be.conv_dgrad(gyc, be.pack_weights(wt.cuda(), 1, dtype), g, impl=IMPL_SIMT)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, _ref_w(wt, dtype), gy.float(), stride=s, padding=p)
    assert rel(dx, dx_ref) < tol
    dw, db = be.conv_wgrad(xc, gyc, g, impl=IMPL_SIMT)
    dw_ref = torch.nn.grad.conv2d_weight(x.float(), wt.shape, gy.float(), stride=s, padding=p)
    assert rel(dw, dw_ref) < max(tol, 2e-5) and rel(db, gy.float().sum((0, 2, 3))) < max(tol, 2e-5)


THIN_CASES = [
    (2, 3, 40, 40, 64, 3, 1, 1),      # RGB -> 64 (conv1, MSB.conv1, D.model.0, VGG.0)
    (2, 3, 30, 31, 64, 1, 1, 0),      # MSB 1x1 from RGB
    (2, 64, 40, 45, 3, 3, 1, 1),      # conv3: 64 -> RGB
    (1, 512, 14, 14, 1, 3, 1, 1),     # D.model.25: 512 -> 1
    (1, 64, 216, 216, 3, 3, 1, 1),    # conv3 at the real output size
    (1, 3, 216, 216, 64, 3, 1, 1),    # D.model.0 at the real input size
]


@pytest.mark.parametrize("case", THIN_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_thin_channel_paths(be, case, dtype):
    """direct kernels for layers with <= 4 channels on one side (auto dispatch) vs torch CPU"""
    n, cin, h, w, cout, k, s, p = case
    x, wt, b = _mk(n, cin, h, w, cout, k, dtype, seed=cin * 3 + cout)
    g = conv_geom(x.shape, wt.shape, s, p)
    tol = 1e-5 if dtype == torch.float32 else 4e-3
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    y = be.conv_fwd(xc, be.pack_weights(wt.cuda(), 0, dtype), b.cuda(), None, g, ACT_LRELU, 0.01, out_dtype=torch.float32)
    y_ref = F.leaky_relu(F.conv2d(x.float(), _ref_w(wt, dtype), b, stride=s, padding=p), 0.01)
    assert y.shape == y_ref.shape and rel(y, y_ref) < 1e-5
    y2 = be.conv_fwd(xc, be.pack_weights(wt.cuda(), 0, dtype), b.cuda(), None, g, ACT_LRELU, 0.01)
    assert rel(y2, y_ref) < tol
    gy = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(1)).to(dtype)
    gyc = gy.cuda().contiguous(memory_format=torch.channels_last)
    dx = be.conv_dgrad(gyc, be.pack_weights(wt.cuda(), 1, dtype), g)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, _ref_w(wt, dtype), gy.float(), stride=s, padding=p)
    assert rel(dx, dx_ref) < tol
    dw, db = be.conv_wgrad(xc, gyc, g)
    dw_ref = torch.nn.grad.conv2d_weight(x.float(), wt.shape, gy.float(), stride=s, padding=p)
    assert rel(dw, dw_ref) < 5e-5 and rel(db, gy.float().sum((0, 2, 3))) < 5e-5


def test_simt_residual_shuffle_fp32out(be):
    n, cin, h, w, cout, k = 2, 64, 9, 9, 256, 3
    x, wt, b = _mk(n, cin, h, w, cout, k, torch.bfloat16, seed=3)
    g = conv_geom(x.shape, wt.shape, 1, 1)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    y = be.conv_fwd(xc, be.pack_weights(wt.cuda(), 0, torch.bfloat16, 2), b.cuda(), None, g, ACT_LRELU, 0.01, shuffle_r=2, impl=IMPL_SIMT)
    y_ref = F.leaky_relu(F.pixel_shuffle(F.conv2d(x.float(), wt.bfloat16().float(), b, padding=1), 2), 0.01)
    assert y.shape == y_ref.shape and rel(y, y_ref) < 4e-3
    res = torch.randn(n, cout, h, w)
    y2 = be.conv_fwd(xc, be.pack_weights(wt.cuda(), 0, torch.bfloat16), b.cuda(), res.cuda().contiguous(memory_format=torch.channels_last),
                     g, out_dtype=torch.float32, impl=IMPL_SIMT)
    assert rel(y2, F.conv2d(x.float(), wt.bfloat16().float(), b, padding=1) + res) < 1e-5


TC_CASES = [
    # n, cin, h, w, cout, k, stride, pad   (x4 generator / D / VGG shapes, incl. non-tile-multiple maps)
    (3, 64, 54, 54, 256, 3, 1, 1),     # K1  RAB.conv1
    (3, 256, 54, 54, 64, 3, 1, 1),     # K2  RAB.conv2
    (2, 64, 27, 27, 64, 1, 1, 0),      # K4  1x1
    (1, 128, 14, 14, 128, 3, 1, 1),    # D / VGG mid layers, tiny map (196 pixels -> 2 tiles, ragged tail)
    (1, 64, 24, 24, 576, 3, 1, 1),     # x3/x9 head, BLOCK_N = 192
    (1, 64, 216, 216, 64, 3, 1, 1),    # 216^2 layers (VGG.2)
    (2, 512, 7, 7, 512, 3, 1, 1),      # D.22-like, K = 4608
    (1, 64, 5, 5, 64, 3, 1, 1),        # fewer pixels than one tile
]


@pytest.mark.parametrize("case", TC_CASES)
def test_tcgen05_fwd_matches_cpu_and_simt(be, case):
    n, cin, h, w, cout, k, s, p = case
    x, wt, b = _mk(n, cin, h, w, cout, k, torch.bfloat16, seed=cin * 7 + cout)
    g = conv_geom(x.shape, wt.shape, s, p)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    wp = be.pack_weights(wt.cuda(), 0, torch.bfloat16)
    y = be.conv_fwd(xc, wp, b.cuda(), None, g, ACT_LRELU, 0.2, impl=IMPL_TCGEN05)
    torch.cuda.synchronize()
    y_ref = F.leaky_relu(F.conv2d(x.float(), wt.bfloat16().float(), b, stride=s, padding=p), 0.2)
    assert y.shape == y_ref.shape
    assert rel(y, y_ref) < 4e-3
    y_simt = be.conv_fwd(xc, wp, b.cuda(), None, g, ACT_LRELU, 0.2, impl=IMPL_SIMT)
    assert rel(y, y_simt) < 4e-3
    # fp32 output + residual epilogue
    res = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(2))
    y2 = be.conv_fwd(xc, wp, b.cuda(), res.cuda().contiguous(memory_format=torch.channels_last), g, out_dtype=torch.float32,
                     impl=IMPL_TCGEN05)
    assert rel(y2, F.conv2d(x.float(), wt.bfloat16().float(), b, stride=s, padding=p) + res) < 2e-5


@pytest.mark.parametrize("case", TC_CASES[:5] + [(1, 64, 216, 216, 64, 3, 2, 1), (2, 128, 27, 27, 128, 3, 2, 1),
                                                 (1, 256, 54, 54, 256, 3, 2, 1), (1, 512, 14, 14, 512, 3, 2, 1)])
def test_tcgen05_dgrad(be, case):
    n, cin, h, w, cout, k, s, p = case
    x, wt, _ = _mk(n, cin, h, w, cout, k, torch.bfloat16, seed=cout)
    g = conv_geom(x.shape, wt.shape, s, p)
    gy = torch.randn(n, cout, g.Ho, g.Wo, generator=torch.Generator().manual_seed(4)).bfloat16()
    gyc = gy.cuda().contiguous(memory_format=torch.channels_last)
    dx = be.conv_dgrad(gyc, be.pack_weights(wt.cuda(), 1, torch.bfloat16), g, impl=IMPL_TCGEN05)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, wt.bfloat16().float(), gy.float(), stride=s, padding=p)
    assert rel(dx, dx_ref) < 4e-3


WGRAD_CASES = TC_CASES + [(2, 64, 216, 216, 64, 3, 2, 1), (1, 256, 54, 54, 256, 3, 2, 1), (2, 128, 27, 27, 256, 3, 1, 1),
                         (1, 64, 30, 30, 64, 1, 1, 0)]


@pytest.mark.parametrize("case", WGRAD_CASES)
def test_tcgen05_wgrad(be, case):
    """MN-major tcgen05 weight gradient (split-K over pixels, fp32 atomics into OIHW) vs torch CPU."""
    n, cin, h, w, cout, k, s, p = case
    x, wt, _ = _mk(n, cin, h, w, cout, k, torch.bfloat16, seed=cin + 3 * cout)
    g = conv_geom(x.shape, wt.shape, s, p)
    gy = (torch.randn(n, cout, g.Ho, g.Wo, generator=torch.Generator().manual_seed(8)) * 0.1).bfloat16()
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    gyc = gy.cuda().contiguous(memory_format=torch.channels_last)
    dw, db = be.conv_wgrad(xc, gyc, g, impl=IMPL_TCGEN05)
    dw_ref = torch.nn.grad.conv2d_weight(x.float(), wt.shape, gy.float(), stride=s, padding=p)
    assert rel(dw, dw_ref) < 2e-5        # bf16 operands are exact products; fp32 accumulation both sides
    assert rel(db, gy.float().sum((0, 2, 3))) < 2e-5
    dw2, _ = be.conv_wgrad(xc, gyc, g, impl=IMPL_SIMT)
    assert rel(dw, dw2) < 2e-5


@pytest.mark.parametrize("r,cout", [(2, 256), (3, 576)])
def test_tcgen05_pixel_shuffle_epilogue(be, r, cout):
    x, wt, b = _mk(2, 64, 12, 12, cout, 3, torch.bfloat16, seed=r)
    g = conv_geom(x.shape, wt.shape, 1, 1)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    y = be.conv_fwd(xc, be.pack_weights(wt.cuda(), 0, torch.bfloat16, r), b.cuda(), None, g, ACT_LRELU, 0.01, shuffle_r=r,
                    impl=IMPL_TCGEN05)
    y_ref = F.leaky_relu(F.pixel_shuffle(F.conv2d(x.float(), wt.bfloat16().float(), b, padding=1), r), 0.01)
    assert y.shape == y_ref.shape and rel(y, y_ref) < 4e-3


@pytest.mark.parametrize("case", [(2, 64, 216, 216, 64, 3, 2, 1), (1, 128, 27, 27, 128, 3, 2, 1), (1, 256, 54, 54, 256, 3, 2, 1)])
def test_tcgen05_stride2_fwd(be, case):
    """Discriminator stride-2 blocks through the im2col TMA traversal stride."""
    n, cin, h, w, cout, k, s, p = case
    x, wt, b = _mk(n, cin, h, w, cout, k, torch.bfloat16, seed=11)
    g = conv_geom(x.shape, wt.shape, s, p)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    y = be.conv_fwd(xc, be.pack_weights(wt.cuda(), 0, torch.bfloat16), b.cuda(), None, g, impl=IMPL_TCGEN05)
    y_ref = F.conv2d(x.float(), wt.bfloat16().float(), b, stride=s, padding=p)
    assert y.shape == y_ref.shape and rel(y, y_ref) < 4e-3


def test_full_size_linearity_and_impl_agreement(be):
    """BASELINE config shape (B=16, 64->256 @54^2): size-independent properties instead of a CPU oracle —
    conv is linear in x (fp32-out epilogue), and the tcgen05 and SIMT paths agree."""
    g_ = torch.Generator().manual_seed(0)
    a = torch.randn(16, 64, 54, 54, generator=g_).bfloat16().cuda().contiguous(memory_format=torch.channels_last)
    b2 = torch.randn(16, 64, 54, 54, generator=g_).bfloat16().cuda().contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(256, 64, 3, 3, generator=g_) / 24).cuda()
    g = conv_geom(a.shape, wt.shape, 1, 1)
    wp = be.pack_weights(wt, 0, torch.bfloat16)
    ya = be.conv_fwd(a, wp, None, None, g, out_dtype=torch.float32, impl=IMPL_TCGEN05)
    yb = be.conv_fwd(b2, wp, None, None, g, out_dtype=torch.float32, impl=IMPL_TCGEN05)
    s = (a.float() + b2.float())
    s_bf = s.bfloat16()
    exact = (s_bf.float() == s).all().item()
    ys = be.conv_fwd(s_bf, wp, None, None, g, out_dtype=torch.float32, impl=IMPL_TCGEN05)
    if exact:
        assert rel(ys, ya + yb) < 1e-5
    y_simt = be.conv_fwd(a, wp, None, None, g, out_dtype=torch.float32, impl=IMPL_SIMT)
    assert rel(ya, y_simt) < 1e-5


HALO_CASES = [
    # n, cin, h, w, cout  (3x3, stride 1, pad 1)
    (2, 64, 54, 54, 256),      # RAB.conv1 (K1): resident weights, 2 rows per tile
    (2, 256, 54, 54, 64),      # RAB.conv2 (K2): streamed weights, 4 channel blocks
    (1, 64, 216, 216, 64),     # 216^2 layers: column strips (2 x 108) with real neighbours in the halo
    (1, 128, 108, 108, 128),   # one padded row per tile
    (3, 256, 27, 27, 512),     # 4 rows per tile, 4 column blocks
    (2, 512, 14, 14, 512),     # 8 rows per tile, last tile ragged
    (1, 64, 5, 7, 64),         # fewer rows than TR
    (2, 64, 130, 33, 128),     # W + 2 <= 129 with 3 rows per tile, ragged height
    (1, 64, 40, 300, 64),      # three column strips
    # 64 output channels from >= 128 input channels (and the input gradients of the mirrored shapes): the three kx taps of a filter
    # row stacked along N, column shift in the epilogue (conv_halo.cu `stack` mode)
    (1, 128, 40, 300, 64),     # three column strips, two channel blocks
    (2, 256, 130, 33, 64),     # three rows per tile, ragged height
    (1, 128, 9, 62, 64),       # TR * TWp = 128: the last output row of a tile reads accumulator rows 126 / 127
    (1, 192, 5, 7, 64),        # fewer rows than TR, three channel blocks
    (3, 64, 27, 27, 128),      # dgrad: 128 -> 64 stacked, odd number of tiles (single-tile items)
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_halo_fwd_dgrad_match_cpu_and_im2col_kernel(be, case):
    """Halo-tile tcgen05 kernel (one TMA box per channel block, nine shifted UMMA views) == torch CPU == the
    im2col-TMA tcgen05 kernel, forward (bias + LeakyReLU, bf16 out; fp32 out + residual) and input gradient."""
    n, cin, h, w, cout = case
    x, wt, b = _mk(n, cin, h, w, cout, 3, torch.bfloat16, seed=cin * 3 + cout + h)
    g = conv_geom(x.shape, wt.shape, 1, 1)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    wp = be.pack_weights(wt.cuda(), 0, torch.bfloat16)
    y = be.conv_fwd(xc, wp, b.cuda(), None, g, ACT_LRELU, 0.2, impl=IMPL_HALO)
    torch.cuda.synchronize()
    y_ref = F.leaky_relu(F.conv2d(x.float(), wt.bfloat16().float(), b, padding=1), 0.2)
    assert y.shape == y_ref.shape
    assert rel(y, y_ref) < 4e-3
    y_tc = be.conv_fwd(xc, wp, b.cuda(), None, g, ACT_LRELU, 0.2, impl=IMPL_TCGEN05)
    assert rel(y, y_tc) < 1e-3          # same bf16 products; fp32 accumulation order differs (channel-block major vs tap major)
    res = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(2))
    y2 = be.conv_fwd(xc, wp, b.cuda(), res.cuda().contiguous(memory_format=torch.channels_last), g, out_dtype=torch.float32,
                     impl=IMPL_HALO)
    assert rel(y2, F.conv2d(x.float(), wt.bfloat16().float(), b, padding=1) + res) < 2e-5
    gy = torch.randn(n, cout, h, w, generator=torch.Generator().manual_seed(4)).bfloat16()
    gyc = gy.cuda().contiguous(memory_format=torch.channels_last)
    dx = be.conv_dgrad(gyc, be.pack_weights(wt.cuda(), 1, torch.bfloat16), g, impl=IMPL_HALO)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, wt.bfloat16().float(), gy.float(), stride=1, padding=1)
    assert rel(dx, dx_ref) < 4e-3


@pytest.mark.parametrize("case", [(2, 256, 54, 54, 64), (1, 128, 31, 45, 64)])
def test_halo_stack_mode_residual_and_plain_mode_agree(be, case):
    """stack mode (N = 192 instructions + epilogue column shift) against the nine-tap mode of the same kernel (SR_HALO_STACK 0)
    and torch CPU, with a bf16 residual in the epilogue."""
    from sradsgan_b200 import _lib
    n, cin, h, w, cout = case
    x, wt, b = _mk(n, cin, h, w, cout, 3, torch.bfloat16, seed=7)
    g = conv_geom(x.shape, wt.shape, 1, 1)
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
    res = torch.randn(n, cout, h, w, generator=torch.Generator().manual_seed(3)).bfloat16()
    wp = be.pack_weights(wt.cuda(), 0, torch.bfloat16)
    lib = _lib.load()
    try:
        lib.sr_set_option(b"SR_HALO_STACK", 0)
        y_plain = be.conv_fwd(cl(x), wp, b.cuda(), cl(res), g, ACT_NONE, 0.0, impl=IMPL_HALO)
        torch.cuda.synchronize()
    finally:
        lib.sr_set_option(b"SR_HALO_STACK", 1)
    y = be.conv_fwd(cl(x), wp, b.cuda(), cl(res), g, ACT_NONE, 0.0, impl=IMPL_HALO)
    y_ref = F.conv2d(x.float(), wt.bfloat16().float(), b, padding=1) + res.float()
    assert rel(y, y_ref) < 4e-3 and rel(y_plain, y_ref) < 4e-3
    assert rel(y, y_plain) < 3e-3       # same products, different fp32 summation order, one bf16 rounding each


@pytest.mark.parametrize("r,cout", [(2, 256), (3, 576)])
def test_halo_pixel_shuffle_epilogue(be, r, cout):
    x, wt, b = _mk(2, 64, 24, 24, cout, 3, torch.bfloat16, seed=r)
    g = conv_geom(x.shape, wt.shape, 1, 1)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    y = be.conv_fwd(xc, be.pack_weights(wt.cuda(), 0, torch.bfloat16, r), b.cuda(), None, g, ACT_LRELU, 0.01, shuffle_r=r,
                    impl=IMPL_HALO)
    y_ref = F.leaky_relu(F.pixel_shuffle(F.conv2d(x.float(), wt.bfloat16().float(), b, padding=1), r), 0.01)
    assert y.shape == y_ref.shape and rel(y, y_ref) < 4e-3


@pytest.mark.parametrize("case", [(2, 64, 54, 54, 256), (1, 256, 27, 27, 64), (1, 64, 20, 20, 64)])
@pytest.mark.parametrize("act", [ACT_LRELU, ACT_RELU])
def test_dgrad_with_fused_activation_derivative(be, case, act):
    """sr_conv2d_dgrad_act: dx = dgrad(dy) * act'(y_prev) in the halo kernel's epilogue == dgrad followed by the mask."""
    from oracle import ops_emu
    n, cin, h, w, cout = case
    x, wt, _ = _mk(n, cin, h, w, cout, 3, torch.bfloat16, seed=cin + cout)
    g = conv_geom(x.shape, wt.shape, 1, 1)
    gen = torch.Generator().manual_seed(11)
    gy = torch.randn(n, cout, h, w, generator=gen).bfloat16()
    y_prev = torch.randn(n, cin, h, w, generator=gen).bfloat16()
    wt_p = be.pack_weights(wt.cuda(), 1, torch.bfloat16)
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
    got = be.conv_dgrad_act(cl(gy), wt_p, g, cl(y_prev), act, 0.2)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, wt.bfloat16().float(), gy.float(), stride=1, padding=1)
    ref = torch.where(y_prev.float() > 0, dx_ref, dx_ref * (0.2 if act == ACT_LRELU else 0.0))
    assert rel(got, ref) < 4e-3
    emu = ops_emu.EmuBackend()
    ref2 = emu.conv_dgrad_act(gy, emu.pack_weights(wt, 1, torch.bfloat16), g, y_prev, act, 0.2)
    assert rel(got, ref2) < 4e-3


@pytest.mark.parametrize("case", [(2, 64, 40, 40, 3), (2, 512, 14, 14, 1), (1, 64, 216, 216, 3), (3, 128, 27, 27, 2)])
def test_halo_thin_output_forward(be, case):
    """Cout <= 4 (conv3 64->3, critic map 512->1) on the halo kernel: one 64-wide column block, only Cout columns stored."""
    n, cin, h, w, cout = case
    x, wt, b = _mk(n, cin, h, w, cout, 3, torch.bfloat16, seed=cin + cout)
    g = conv_geom(x.shape, wt.shape, 1, 1)
    xc = x.cuda().contiguous(memory_format=torch.channels_last)
    wp = be.pack_weights(wt.cuda(), 0, torch.bfloat16)
    for od, tol in ((torch.bfloat16, 4e-3), (torch.float32, 2e-5)):
        y = be.conv_fwd(xc, wp, b.cuda(), None, g, ACT_LRELU, 0.2, out_dtype=od, impl=IMPL_HALO)
        y_ref = F.leaky_relu(F.conv2d(x.float(), wt.bfloat16().float(), b, padding=1), 0.2)
        assert y.shape == y_ref.shape and rel(y, y_ref) < tol, (od, rel(y, y_ref))


@pytest.mark.parametrize("case", [(2, 3, 40, 40, 64), (1, 3, 216, 216, 64), (2, 1, 14, 14, 512)])
def test_halo_thin_input_gradient(be, case):
    """dgrad of a conv with Cin <= 4 (RGB-side layers): thin OUTPUT of the transposed convolution."""
    n, cin, h, w, cout = case
    x, wt, _ = _mk(n, cin, h, w, cout, 3, torch.bfloat16, seed=cin + cout)
    g = conv_geom(x.shape, wt.shape, 1, 1)
    gy = torch.randn(n, cout, h, w, generator=torch.Generator().manual_seed(4)).bfloat16()
    gyc = gy.cuda().contiguous(memory_format=torch.channels_last)
    dx = be.conv_dgrad(gyc, be.pack_weights(wt.cuda(), 1, torch.bfloat16), g, impl=IMPL_HALO)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, wt.bfloat16().float(), gy.float(), stride=1, padding=1)
    assert dx.shape == dx_ref.shape and rel(dx, dx_ref) < 4e-3


def test_batched_weight_packing_matches_single_packs(be):
    """ops.PackPlan: ONE sr_pack_weights_batched launch rewrites every cached packed operand of a set of parameters
    (both layouts, PixelShuffle row permutation, ragged sizes) exactly as the per-weight kernel does."""
    from sradsgan_b200 import ops
    g = torch.Generator().manual_seed(11)
    shapes = [(256, 64, 3, 3, 2), (64, 256, 3, 3, 0), (64, 3, 3, 3, 0), (3, 64, 3, 3, 0), (64, 64, 1, 1, 0), (1, 2, 7, 7, 0), (576, 64, 3, 3, 3)]
    params = [torch.nn.Parameter(torch.randn(co, ci, k, k2, generator=g).cuda()) for co, ci, k, k2, _ in shapes]
    for w, sh in zip(params, shapes):
        ops.packed(w, 0, torch.bfloat16, sh[4])
        ops.packed(w, 1, torch.bfloat16, 0)
    plan = ops.PackPlan(params)
    assert len(plan.entries) == 2 * len(shapes) and plan.valid()
    with torch.no_grad():
        for w in params:
            w.data.mul_(-1.7).add_(0.3)          # in-place change of the masters (what the fused Adam kernel does)
    plan.repack()
    torch.cuda.synchronize()
    for w, sh in zip(params, shapes):
        for mode, r in ((0, sh[4]), (1, 0)):
            got = ops.packed(w, mode, torch.bfloat16, r)            # cache hit: the plan's persistent tensor
            want = be.pack_weights(w, mode, torch.bfloat16, r)
            assert got.data_ptr() in [e[2].data_ptr() for e in plan.entries]
            assert torch.equal(got, want), (sh, mode)
