"""2+ ranks (scripts/run_ranks.sh): the data-parallel graph replay with the generator's all-reduce + Adam overlapped with the D
phase (SR_DP_OVERLAP=1, default) must leave exactly the parameters of the serial order (SR_DP_OVERLAP=0) after the same steps on
the same data; prints the time per step of both."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from sradsgan_b200.model.sradsgan import SRADSGAN

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
else:
    # one GPU: the four-segment replay with a no-op all-reduce — Adam_G still runs on the communication stream next to the D phase
    os.environ["SR_DP_FORCE_SEGMENTS"] = "1"


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
B = int(os.environ.get("SR_BATCH", "16"))
STEPS = int(os.environ.get("SR_STEPS", "12"))
g = torch.Generator().manual_seed(1234 + rank)
hr = torch.rand(B, 3, 216, 216, generator=g).cuda()
lr = torch.nn.functional.interpolate(hr, size=54, mode="bicubic", align_corners=False).clamp(0, 1)
res = {}
for tag, overlap in (("0", "0"), ("0b", "0"), ("1", "1")):      # the serial order twice: the run-to-run spread of the replay itself
    os.environ["SR_DP_OVERLAP"] = overlap
    torch.manual_seed(0)
    np.random.seed(7)
    net = SRADSGAN(bench.trainer_args(batch_size=B))
    net.build(init=True)
    for _ in range(3):
        net.graphed_step(lr, hr)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        out = net.graphed_step(lr, hr)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / STEPS], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[tag] = (t.item(), net.optimizer_G.flat_param.clone(), net.optimizer_D.flat_param.clone(), float(out["loss_G"]), float(out["loss_D"]))
    del net
    torch.cuda.empty_cache()
same_g = torch.equal(res["0"][1], res["1"][1])
same_d = torch.equal(res["0"][2], res["1"][2])
dg = (res["0"][1] - res["1"][1]).abs().max().item()
dd = (res["0"][2] - res["1"][2]).abs().max().item()
# every rank must also hold the same replica
chk = torch.stack([res["1"][1].double().sum(), res["1"][2].double().sum()])
lo, hi = chk.clone(), chk.clone()
if world > 1:
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print("world %d batch %d/GPU: serial %.3f ms/step, overlapped %.3f ms/step" % (world, B, res["0"][0], res["1"][0]))
    print("parameters after %d steps identical: G %s (max |d| %.3g)  D %s (max |d| %.3g); losses serial %.6f / %.6f overlapped %.6f / %.6f" % (
        STEPS + 3, same_g, dg, same_d, dd, res["0"][3], res["0"][4], res["1"][3], res["1"][4]))
    print("serial vs serial (run-to-run spread of the replay): G max |d| %.3g  D max |d| %.3g" % (
        (res["0"][1] - res["0b"][1]).abs().max().item(), (res["0"][2] - res["0b"][2]).abs().max().item()))
    print("replicas agree across ranks:", bool(torch.equal(lo, hi)))
if world > 1:
    dist.destroy_process_group()
