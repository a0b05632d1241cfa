"""Role timeline of conv_halo_kernel at the RAB shapes (diagnostics build: python __graft_entry__.py --probes, run with
SR_LIB_PATH=build/probes/libsradsgan_b200.so).  Prints, for a few CTAs, the SM-clock stamps of the producer / MMA issuers /
epilogue groups relative to the CTA's first instruction — the picture behind the `tensor pipe active` percentage."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sradsgan_b200 import _lib
from sradsgan_b200._lib import ACT_LRELU, ACT_NONE, conv_geom

SLOTS = 128
B = int(os.environ.get("SR_BATCH", "16"))
H = int(os.environ.get("SR_H", "54"))
DBG = [int(v) for v in os.environ.get("SR_DBG", "0").split(",")]
CTAS = [int(v) for v in os.environ.get("SR_CTAS", "0,73").split(",")]


def show(tr, ctas):
    for c in ctas:
        t = tr[c]
        t0 = t[0]
        rel = lambda s: (t[s] - t0) if t[s] else None
        print("  CTA %3d: setup done %s, end %s" % (c, rel(1), rel(2)))
        items = [it for it in range(8) if t[8 + it * 2] or t[8 + it * 2 + 1]]
        for it in items:
            line = "    item %d: load issued %s/%s |" % (it, rel(8 + it * 2), rel(8 + it * 2 + 1))
            for w in range(2):
                b = 32 + w * 24 + it * 3
                line += " iss%d acc_free %s a_landed %s committed %s |" % (w, rel(b), rel(b + 1), rel(b + 2))
            for g in range(2):
                b = 80 + g * 16 + it * 2
                line += " epi%d full %s done %s |" % (g, rel(b), rel(b + 1))
            print(line)


def main():
    be = _lib.backend()
    be.device_check()
    lib = _lib.load()
    dt = torch.bfloat16
    tr = torch.zeros(148 * SLOTS, dtype=torch.int64, device="cuda")
    for name, cin, cout, act in (("K1 64->256", 64, 256, ACT_LRELU), ("K2 256->64", 256, 64, ACT_NONE)):
        x = torch.randn(B, cin, H, H, device="cuda").to(dt).contiguous(memory_format=torch.channels_last)
        w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
        b = torch.randn(cout, device="cuda")
        g = conv_geom(x.shape, w.shape, 1, 1)
        wp = be.pack_weights(w, 0, dt, 0)
        wt = be.pack_weights(w, 1, dt, 0)
        dy = torch.randn(B, cout, H, H, device="cuda").to(dt).contiguous(memory_format=torch.channels_last)
        for what, fn in (("fwd", lambda: be.conv_fwd(x, wp, b, None, g, act, 0.2, 0)), ("dgrad", lambda: be.conv_dgrad(dy, wt, g))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            for dbg in DBG:
                tr.zero_()
                lib.sr_debug_halo_trace(ctypes.c_void_p(tr.data_ptr()), dbg)
                fn()
                torch.cuda.synchronize()
                lib.sr_debug_halo_trace(None, 0)
                t = tr.view(148, SLOTS).cpu().tolist()
                ends = [r[2] - r[0] for r in t if r[2]]
                print("%s %s dbg=%d: CTA lifetime cycles min %d median %d max %d" % (name, what, dbg, min(ends), sorted(ends)[len(ends) // 2], max(ends)))
                show(t, CTAS)
            # five back-to-back launches, each with its own stamp buffer: window of CTA activity (global timer) vs launch period
            bufs = [torch.zeros(148 * SLOTS, dtype=torch.int64, device="cuda") for _ in range(5)]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for bf in bufs:
                lib.sr_debug_halo_trace(ctypes.c_void_p(bf.data_ptr()), 0)
                fn()
            e1.record()
            torch.cuda.synchronize()
            lib.sr_debug_halo_trace(None, 0)
            ws = []
            for bf in bufs:
                t = bf.view(148, SLOTS).cpu()
                st, en = t[:, 3][t[:, 3] > 0], t[:, 4][t[:, 4] > 0]
                ws.append((int(st.min()), int(st.max()), int(en.min()), int(en.max())))
            base = ws[0][0]
            print("  5 launches back to back: %.1f us per launch by events; per launch [first CTA start, last CTA start, first end, last end] ns:" % (e0.elapsed_time(e1) * 1e3 / 5))
            for w in ws:
                print("    ", [v - base for v in w])


if __name__ == "__main__":
    main()
