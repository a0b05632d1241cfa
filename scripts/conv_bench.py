"""Per-shape timing of the convolution kernels (fwd / dgrad / wgrad) at the shapes of the x4 B=16 training step.
CUDA events on the launching stream, 3 warm-up + N timed launches of the SAME call (inputs stay L2-warm, as
they mostly are inside the step where the producer kernel just wrote them).  Output: one line per shape and
a JSON summary (gpurun_out/conv_bench.json) used to pick the kernel to work on next."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sradsgan_b200 import _lib
from sradsgan_b200._lib import ACT_LRELU, ACT_NONE, conv_geom

B = int(os.environ.get("SR_BATCH", "16"))
ITERS = int(os.environ.get("SR_ITERS", "20"))
ONLY = os.environ.get("SR_ONLY", "")
IMPL = int(os.environ.get("SR_IMPL", "0"))              # 0 auto, 2 im2col tcgen05, 3 halo tcgen05
PROFILE = os.environ.get("SR_PROFILE", "") == "1"     # one launch per kernel, no warm-up (for ncu --set full)

# name, Cin, Cout, k, stride, H(in), act, shuffle_r, calls per training step (fwd, dgrad, wgrad)
SHAPES = [
    ("G.K1 64->256 3x3 @54", 64, 256, 3, 1, 54, ACT_LRELU, 0, (36, 36, 36)),
    ("G.K2 256->64 3x3 @54", 256, 64, 3, 1, 54, ACT_NONE, 0, (36, 36, 36)),
    ("G.up 64->256 3x3 @54 ps2", 64, 256, 3, 1, 54, ACT_LRELU, 2, (1, 1, 1)),
    ("G.up 64->256 3x3 @108 ps2", 64, 256, 3, 1, 108, ACT_LRELU, 2, (1, 1, 1)),
    ("G.1x1 64->64 @54", 64, 64, 1, 1, 54, ACT_NONE, 0, (2, 2, 2)),
    ("G.msb 192->64 1x1 @54", 192, 64, 1, 1, 54, ACT_LRELU, 0, (1, 1, 1)),
    ("V.1 64->64 3x3 @216", 64, 64, 3, 1, 216, ACT_NONE, 0, (2, 1, 0)),
    ("V.2 64->128 3x3 @108", 64, 128, 3, 1, 108, ACT_NONE, 0, (2, 1, 0)),
    ("V.3 128->128 3x3 @108", 128, 128, 3, 1, 108, ACT_NONE, 0, (2, 1, 0)),
    ("V.4 128->256 3x3 @54", 128, 256, 3, 1, 54, ACT_NONE, 0, (2, 1, 0)),
    ("D.2 64->64 s2 @216", 64, 64, 3, 2, 216, ACT_NONE, 0, (6, 5, 5)),
    ("D.3 64->128 @108", 64, 128, 3, 1, 108, ACT_NONE, 0, (6, 5, 5)),
    ("D.4 128->128 s2 @108", 128, 128, 3, 2, 108, ACT_NONE, 0, (6, 5, 5)),
    ("D.5 128->256 @54", 128, 256, 3, 1, 54, ACT_NONE, 0, (6, 5, 5)),
    ("D.6 256->256 s2 @54", 256, 256, 3, 2, 54, ACT_NONE, 0, (6, 5, 5)),
    ("D.7 256->512 @27", 256, 512, 3, 1, 27, ACT_NONE, 0, (6, 5, 5)),
    ("D.8 512->512 s2 @27", 512, 512, 3, 2, 27, ACT_NONE, 0, (6, 5, 5)),
]


def timeit(fn):
    if PROFILE:
        fn()
        torch.cuda.synchronize()
        return 1.0
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ITERS):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS * 1e-3


def main():
    be = _lib.backend()
    be.device_check()
    dt = torch.bfloat16
    out = []
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    for name, cin, cout, k, s, h, act, r, calls in SHAPES:
        if ONLY and ONLY not in name:
            continue
        pad = k // 2
        x = torch.randn(B, cin, h, h, device="cuda").to(dt).contiguous(memory_format=torch.channels_last)
        w = torch.randn(cout, cin, k, k, device="cuda") * 0.05
        b = torch.randn(cout, device="cuda")
        g = conv_geom(x.shape, w.shape, s, pad)
        flops = 2.0 * g.N * g.Ho * g.Wo * cout * cin * k * k
        wp = be.pack_weights(w, 0, dt, r)
        wt = be.pack_weights(w, 1, dt, 0)
        dy = torch.randn(B, cout, g.Ho, g.Wo, device="cuda").to(dt).contiguous(memory_format=torch.channels_last)
        t_f = timeit(lambda: be.conv_fwd(x, wp, b, None, g, act, 0.2, r, impl=IMPL if (s == 1 and k == 3) else 0))
        t_d = timeit(lambda: be.conv_dgrad(dy, wt, g, impl=IMPL if (s == 1 and k == 3) else 0))
        t_w = timeit(lambda: be.conv_wgrad(x, dy, g, want_bias=True))
        rec = {"shape": name, "gflop": flops / 1e9, "fwd_us": t_f * 1e6, "dgrad_us": t_d * 1e6, "wgrad_us": t_w * 1e6,
               "fwd_tflops": flops / t_f / 1e12, "dgrad_tflops": flops / t_d / 1e12, "wgrad_tflops": flops / t_w / 1e12,
               "calls_per_step": calls}
        out.append(rec)
        tot["fwd"] += t_f * calls[0]; tot["dgrad"] += t_d * calls[1]; tot["wgrad"] += t_w * calls[2]
        print("%-28s %7.2f GF | fwd %7.1f us %6.0f TF | dgrad %7.1f us %6.0f TF | wgrad %7.1f us %6.0f TF | step ms %.2f/%.2f/%.2f" % (
            name, flops / 1e9, t_f * 1e6, rec["fwd_tflops"], t_d * 1e6, rec["dgrad_tflops"], t_w * 1e6, rec["wgrad_tflops"],
            t_f * calls[0] * 1e3, t_d * calls[1] * 1e3, t_w * calls[2] * 1e3), flush=True)
        del x, w, dy
    print("per-step totals (ms): fwd %.2f dgrad %.2f wgrad %.2f" % (tot["fwd"] * 1e3, tot["dgrad"] * 1e3, tot["wgrad"] * 1e3))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"batch": B, "shapes": out, "step_ms": {k: v * 1e3 for k, v in tot.items()}}, open("gpurun_out/conv_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
