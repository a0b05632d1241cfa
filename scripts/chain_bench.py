"""BASELINE.json configs[2]: chain training x2 -> x3 -> x4 (each stage warm-started from the previous one, reference
model/sradsgan.py:716-721), data parallel, batch 16 per GPU, HR 216^2 (LR 108 / 72 / 54), bf16, CUDA-graph replay.
Launch with scripts/run_ranks.sh N LIMIT scripts/chain_bench.py (or single process).  Rank 0 prints one JSON line per stage:
HR images/s over all ranks (CUDA events, max over ranks, barrier on both sides), plus the warm-start bookkeeping."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from sradsgan_b200.model.sradsgan import SRADSGAN

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(os.environ.get("SR_BATCH", "16"))
STEPS = int(os.environ.get("SR_STEPS", "16"))
WARM = int(os.environ.get("SR_WARMUP", "3"))
HR = 216


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


prev = None
for scale in (2, 3, 4):
    net = SRADSGAN(bench.trainer_args(batch_size=B, scale_factor=scale))
    net.build(init=True)
    loaded = skipped = 0
    if prev is not None:
        lg, sg = SRADSGAN.warm_start(net.generator, prev[0])
        ld, sd_ = SRADSGAN.warm_start(net.discriminator, prev[1])
        loaded, skipped = len(lg) + len(ld), len(sg) + len(sd_)
    g = torch.Generator().manual_seed(1234 + rank + 100 * scale)
    hr = torch.rand(B, 3, HR, HR, generator=g).cuda()
    lr = torch.nn.functional.interpolate(hr, size=HR // scale, mode="bicubic", align_corners=False).clamp(0, 1)
    for _ in range(WARM):
        out = net.graphed_step(lr, hr)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        out = net.graphed_step(lr, hr)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / STEPS
    finite = bool(torch.isfinite(out["loss_G"]).item() and torch.isfinite(out["loss_D"]).item())
    if rank == 0:
        print(json.dumps({"stage": "x%d" % scale, "n_gpus": world, "global_batch": B * world, "lr_size": HR // scale, "ms_per_step": ms,
                          "value": B * world / (ms * 1e-3), "unit": "HR images/s", "steps": STEPS, "warmup": WARM, "loss_finite": finite,
                          "warm_start": {"loaded": loaded, "reinitialised": skipped}}), flush=True)
    prev = ({k: v.detach().clone() for k, v in net.generator.state_dict().items()},
            {k: v.detach().clone() for k, v in net.discriminator.state_dict().items()})
    del net
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
