"""Timing of the fused local-attention chain (fwd / bwd) at the x4 B=16 shape, CUDA events, L2-warm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sradsgan_b200 import _lib
be = _lib.backend()
N, H, W = int(os.environ.get("SR_BATCH", "16")), 54, 54
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(N, 64, H, W, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
t = torch.randn(N, 64, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
fc1 = torch.randn(4, 64, 1, 1, device="cuda") * 0.3; fc2 = torch.randn(64, 4, 1, 1, device="cuda") * 0.3
w7 = torch.randn(1, 2, 7, 7, device="cuda") * 0.2; Wm = torch.randn(64, 64, 1, 1, device="cuda") * 0.125; b = torch.randn(64, device="cuda") * 0.1
gz32 = torch.randn_like(t); gz16 = torch.randn_like(x)
def timeit(fn, it=20):
    """GPU time per call with the host out of the picture: `it` calls captured into one CUDA graph, replayed"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(it): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * it) * 1e3
z32, z16, sv = be.la_chain_fwd(x, t, fc1, fc2, w7, Wm, b)
print("la_chain fwd %.1f us   bwd %.1f us   (ideal traffic fwd 36 MB, bwd ~42 MB)" % (
    timeit(lambda: be.la_chain_fwd(x, t, fc1, fc2, w7, Wm, b)), timeit(lambda: be.la_chain_bwd(gz32, gz16, x, sv, fc1, fc2, w7, Wm))))
# band path fed by producer partials (what the RAB conv2 epilogue provides) + dense-sampling accumulator
if be.la_band_path(x):
    _, _, _, _, pool = be.la_chain_forward(x, t, fc1, fc2, w7, Wm, b, want_pool=True)
    acc = torch.randn_like(t)
    gacc = torch.randn_like(t)
    f1 = timeit(lambda: be.la_chain_forward(x, t, fc1, fc2, w7, Wm, b, pool=pool))
    f2 = timeit(lambda: be.la_chain_forward(x, t, fc1, fc2, w7, Wm, b, pool=pool, acc=acc, want_pool=True))
    b2 = timeit(lambda: be.la_chain_backward(gz32, gz16, gacc, x, sv, fc1, fc2, w7, Wm))
    print("band path: fwd with producer partials %.1f us; + accumulator + out partials %.1f us; bwd with accumulator gradient %.1f us" % (f1, f2, b2))
    print("  fwd %.0f GB/s, bwd %.0f GB/s of algorithmic traffic (768 / 896 B per pixel)" % (
        N * H * W * 768 / f1 * 1e-3, N * H * W * 896 / timeit(lambda: be.la_chain_bwd(gz32, gz16, x, sv, fc1, fc2, w7, Wm)) * 1e-3))
