"""Per-kernel period of a dependent chain of RAB convolutions inside a CUDA graph (K1 64->256 -> K2 256->64 -> K1 ...), against the CTA
lifetimes of scripts/halo_trace.py: what one launch costs beyond its own blocks under graph replay, with and without programmatic
dependent launch (SR_PDL)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sradsgan_b200 import _lib
from sradsgan_b200._lib import ACT_LRELU, ACT_NONE, conv_geom

B, H, PAIRS, REPS = 16, 54, 12, 20
be = _lib.backend()
be.device_check()
dt = torch.bfloat16
x = torch.randn(B, 64, H, H, device="cuda").to(dt).contiguous(memory_format=torch.channels_last)
w1 = torch.randn(256, 64, 3, 3, device="cuda") * 0.03
w2 = torch.randn(64, 256, 3, 3, device="cuda") * 0.03
b1, b2 = torch.zeros(256, device="cuda"), torch.zeros(64, device="cuda")
g1, g2 = conv_geom(x.shape, w1.shape, 1, 1), conv_geom((B, 256, H, H), w2.shape, 1, 1)
p1, p2 = be.pack_weights(w1, 0, dt, 0), be.pack_weights(w2, 0, dt, 0)


def chain():
    t = x
    for _ in range(PAIRS):
        y = be.conv_fwd(t, p1, b1, None, g1, ACT_LRELU, 0.2, 0)
        t = be.conv_fwd(y, p2, b2, None, g2, ACT_NONE, 0.0, 0)
    return t


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        chain()
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    out = chain()
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(REPS):
    graph.replay()
e1.record()
torch.cuda.synchronize()
per = e0.elapsed_time(e1) * 1e3 / (REPS * PAIRS * 2)
print("SR_PDL=%s: %d dependent conv launches per graph, %.2f us per launch (K1 + K2 pair %.2f us)" % (os.environ.get("SR_PDL", "0"), PAIRS * 2, per, 2 * per))
