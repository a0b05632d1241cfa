"""CPU, world_size 2, gloo: the data-parallel plumbing (flat-bucket all-reduce, 1/world folded into Adam, optional
chunked overlap hooks incl. gradients that kernels accumulate in place) around the emulated backend."""
import os
import socket
import sys
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(**kw):
    base = dict(model_name="SRADSGAN", train_dataset=[], test_dataset=[], crop_size=32, test_crop_size=32, hr_height=32,
                hr_width=32, num_threads=0, num_channels=3, scale_factor=4, epoch=0, num_epochs=1, save_epochs=1,
                batch_size=2, test_batch_size=1, lr=2e-4, b1=0.9, b2=0.999, data_dir="", root_dir="", save_dir="/tmp/sr_dp",
                gpu_mode=True, n_cpu=0, sample_interval=1000, clip_value=0.01, lambda_gp=10, gp=True, penalty_type="LS",
                grad_penalty_Lp_norm="L2", relativeGan=False, loss_Lp_norm="L1", weight_gan=1e-3, weight_content=1e-2,
                max_train_samples=10, precision="fp32", seed=0)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _worker(rank, world, port, overlap, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SR_DP_OVERLAP="1" if overlap else "0")
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ops_emu, sradsgan_oracle as O
        from sradsgan_b200 import _lib, ops
        from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup, SRADSGAN
        _lib.set_backend(ops_emu.EmuBackend())
        ops.set_precision("fp32")
        net = SRADSGAN(_args())
        net.new_generator = lambda: GeneratorResNet(ResGroup, n_residual_blocks=2, n_basic_blocks=1, upscale_factor=4)
        torch.manual_seed(100 + rank)             # different initial weights per rank: build() must broadcast rank 0's
        net.seed = 100 + rank
        net.build(init=True)
        p0 = [net.optimizer_G.flat_param.clone(), net.optimizer_D.flat_param.clone()]
        lr, hr = O.synthetic_batch(2, 4, 32, seed=7 + rank)
        net._alpha_override = torch.full((2, 1, 1, 1), 0.25 + 0.5 * rank)
        # G phase: the reduced bucket equals the mean of the per-rank gradients
        net._g_phase(lr, hr)
        local = net.optimizer_G.flat_grad.clone() if not overlap else None
        scale = net.reducer_G.finish()
        reduced = net.optimizer_G.flat_grad.clone() * scale
        net.optimizer_G.zero_grad()
        net.reducer_G = type(net.reducer_G)(net.optimizer_G, overlap=False) if overlap else net.reducer_G
        if overlap:                               # recompute the local gradients without any hook-driven reduction
            from sradsgan_b200 import dp
            keep = net.reducer_G
            net.reducer_G = dp.NullReducer(world)
            net._g_phase(lr, hr)
            local = net.optimizer_G.flat_grad.clone()
            net.reducer_G = keep
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        mean = sum(gathered) / world
        err = ((reduced - mean).norm() / mean.norm()).item()
        # a full step afterwards: replicas stay identical
        net.train_step(lr, hr)
        flat = torch.cat([net.optimizer_G.flat_param, net.optimizer_D.flat_param])
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        init = [torch.empty_like(p0[0]) for _ in range(world)]
        dist.all_gather(init, p0[0])
        q.put((rank, err, (both[0] - both[1]).abs().max().item(), (init[0] - init[1]).abs().max().item(),
               (flat[:p0[0].numel()] - p0[0]).abs().max().item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [False, True])
def test_two_rank_gradient_allreduce_and_replica_consistency(overlap):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, replica_diff, init_diff, moved in res:
        assert err < 1e-5, (rank, err)                 # all-reduced bucket / world == mean of the local gradients
        assert init_diff == 0.0                        # rank 0's initial weights were broadcast
        assert replica_diff == 0.0                     # identical updates on both ranks
        assert moved > 0                               # and the step did change the weights


def _edsr_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import edsr_oracle as E, ops_emu
        from sradsgan_b200 import _lib, ops
        from sradsgan_b200.model.edsr import EDSR
        _lib.set_backend(ops_emu.EmuBackend())
        ops.set_precision("fp32")
        net = EDSR(_args(model_name="EDSR", lr=1e-4))
        net.num_residuals = 1
        net.seed = 200 + rank                     # different initial weights per rank: build() must broadcast rank 0's
        net.build(init=True)
        p0 = net.optimizer_G.flat_param.clone()
        lr, hr = E.synthetic_batch(2, 4, 32, seed=17 + rank)
        net._g_phase(lr, hr)
        local = net.optimizer_G.flat_grad.clone()
        scale = net.reducer_G.finish()
        reduced = net.optimizer_G.flat_grad.clone() * scale
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        mean = sum(gathered) / world
        err = ((reduced - mean).norm() / mean.norm()).item()
        net.train_step(lr, hr)
        flat = net.optimizer_G.flat_param
        both = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        init = [torch.empty_like(p0) for _ in range(world)]
        dist.all_gather(init, p0)
        q.put((rank, err, (both[0] - both[1]).abs().max().item(), (init[0] - init[1]).abs().max().item(), (flat - p0).abs().max().item()))
    finally:
        dist.destroy_process_group()


def test_edsr_two_rank_gradient_allreduce_and_replica_consistency():
    """the same data-parallel plumbing under the EDSR trainer (one network, one bucket)"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_edsr_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, replica_diff, init_diff, moved in res:
        assert err < 1e-5, (rank, err)
        assert init_diff == 0.0 and replica_diff == 0.0 and moved > 0
