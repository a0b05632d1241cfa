"""Per-phase GPU time of one training step (CUDA events between phases, eager launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from sradsgan_b200.model.sradsgan import SRADSGAN

B = int(os.environ.get("SR_BATCH", "16"))
net = SRADSGAN(bench.trainer_args(batch_size=B))
net.build(init=True)
hr = torch.rand(B, 3, 216, 216, device="cuda")
lr = torch.nn.functional.interpolate(hr, size=54, mode="bicubic", align_corners=False).clamp(0, 1)
for _ in range(2):
    net.train_step(lr, hr)
torch.cuda.synchronize()
for rep in range(2):
    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    net._phase_mark = mark
    net.train_step(lr, hr)
    torch.cuda.synchronize()
    net._phase_mark = None
    tot = marks[0][1].elapsed_time(marks[-1][1])
    print("rep", rep, "total %.2f ms" % tot)
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        print("   %-20s %8.2f ms" % (n1, e0.elapsed_time(e1)))
print("max mem GB", torch.cuda.max_memory_allocated() / 2**30)
