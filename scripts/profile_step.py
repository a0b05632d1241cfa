"""Runs ONE profiled training step (after warm-up) between cudaProfilerStart/Stop — for
`ncu --profile-from-start off --metrics gpu__time_duration.sum ...` launch lists (profiles/)."""
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
torch.cuda.cudart().cudaProfilerStart()
net.train_step(lr, hr)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")
