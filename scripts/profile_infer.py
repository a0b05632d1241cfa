"""Runs ONE generator forward of a batch of x9 inference tiles between cudaProfilerStart/Stop — for
`ncu --profile-from-start off --metrics gpu__time_duration.sum ...` launch lists of the inference path (profiles/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup
from sradsgan_b200.utils import weights_init_normal

TILE = int(os.environ.get("SR_TILE", "128"))
TB = int(os.environ.get("SR_TILE_BATCH", "8"))
SCALE = int(os.environ.get("SR_SCALE", "9"))
torch.manual_seed(0)
G = GeneratorResNet(ResGroup, n_residual_blocks=12, n_basic_blocks=3, upscale_factor=SCALE)
G.apply(weights_init_normal)
G.cuda().eval()
x = torch.rand(TB, 3, TILE, TILE, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = G(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = G(x)
    e1.record()
    torch.cuda.synchronize()
    print("forward of %d tiles of %d^2 (x%d): %.2f ms = %.1f output Mpix/s (no overlap accounted)" % (
        TB, TILE, SCALE, e0.elapsed_time(e1), TB * (TILE * SCALE) ** 2 / 1e6 / (e0.elapsed_time(e1) * 1e-3)))
    torch.cuda.cudart().cudaProfilerStart()
    y = G(x)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("profiled one forward", tuple(y.shape))
