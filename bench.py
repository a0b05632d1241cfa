#!/usr/bin/env python
"""bench.py — SRADSGAN x4 full GAN training step (BASELINE.json configs[1]) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one complete training iteration of the reference (model/sradsgan.py:829-892): generator
forward/backward with L1 + VGG19[:12] perceptual + adversarial loss, Adam(G); discriminator real/fake +
WGAN-GP (double backward), Adam(D) + weight clamp — on a synthetic batch of 16 HR 216x216 / LR 54x54
images per GPU, bf16 compute, random-init weights of the exact architecture (no network for checkpoints).
Prints ONE JSON line on rank 0.  `--impl reference` times the reference's CPU implementation of the same
step (the oracle port over the same ATen ops) on the host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_IMG = 364.3e9        # minimal-equivalent x4 training step, SURVEY.md §8d / BASELINE.md §2
BATCH = 16
SCALE = 4
HR = 216


def trainer_args(**kw):
    base = dict(model_name="SRADSGAN", train_dataset=[], test_dataset=[], crop_size=HR, test_crop_size=HR, hr_height=HR,
                hr_width=HR, num_threads=0, num_channels=3, scale_factor=SCALE, epoch=0, num_epochs=1, save_epochs=1,
                batch_size=BATCH, test_batch_size=1, lr=2e-4, b1=0.9, b2=0.999, data_dir="", root_dir="",
                save_dir="/tmp/sradsgan_bench", gpu_mode=True, n_cpu=0, sample_interval=1000, clip_value=0.01, lambda_gp=10,
                gp=True, penalty_type="LS", grad_penalty_Lp_norm="L2", relativeGan=False, loss_Lp_norm="L1", weight_gan=1e-3,
                weight_content=1e-2, max_train_samples=0, precision="bf16", seed=0)
    base.update(kw)
    return types.SimpleNamespace(**base)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step(batch, steps, warmup):
    """The reference's CPU path on all host cores, at the SAME configuration as the GPU arm (x4, batch `batch`, HR 216^2):
    the UNMODIFIED reference modules imported through oracle/ref_shim.py when a reference tree is present (build container;
    `kind` = "reference"), else the oracle port — the same ATen ops, pinned to the reference by tests/golden (`kind` = "port";
    the GPU box has no /root/reference)."""
    import numpy as np
    import torch
    from oracle import ref_shim
    from oracle import sradsgan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Gsd = O.tie_upsampling(O.make_state(O.generator_spec(SCALE), seed=0, init="ref"))
    Dsd = O.make_state(O.discriminator_spec(), seed=1, init="ref")
    Vsd = O.make_state(O.vgg_spec(), seed=2, init="fan")
    lr, hr = O.synthetic_batch(batch, SCALE, HR, seed=1234)
    rs = np.random.RandomState(1234)
    if ref_shim.available():
        from oracle import make_golden as MG
        ref = ref_shim.load_reference()
        G, D, V = MG.build_ref_generator(ref, Gsd, SCALE, 12, 3), MG.build_ref_discriminator(ref, Dsd), MG.build_ref_vgg(Vsd)
        oG = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.9, 0.999))
        oD = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.9, 0.999))
        kind = "reference"
        step = lambda i: MG.ref_train_step(ref, G, D, V, oG, oD, lr, hr, np_seed=1234 + i)
    else:
        st = O.TrainState(Gsd, Dsd, Vsd, SCALE)
        kind = "port"
        step = lambda i: O.train_step(st, lr, hr, torch.from_numpy(rs.random((batch, 1, 1, 1))).float())
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(i)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return {"value": batch / mean, "unit": "HR images/s", "cores": cores, "kind": kind,
            "sample": "%d full G+D training step(s) of the same workload (x4, batch %d, HR 216^2 / LR 54^2) after %d warm-up, fp32, "
                      "%d threads; %.2f s/step" % (steps, batch, warmup, cores, mean)}, mean


def g_forward_bench(batch, net):
    """BASELINE.json configs[0]: SRADSGAN x4 generator forward on a synthetic 54x54 LR batch of 16 -> 216x216, the reference's
    model on the CPU beside this build on the GPU (inputs resident, no_grad, eager launches, CUDA events)."""
    import torch
    from oracle import sradsgan_oracle as O
    lr, _ = O.synthetic_batch(batch, SCALE, HR, seed=1234)
    x = lr.cuda()
    G = net.generator
    with torch.no_grad():
        for _ in range(3):
            G(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            G(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = O.tie_upsampling(O.make_state(O.generator_spec(SCALE), seed=0, init="ref"))
    best = None
    with torch.no_grad():
        for i in range(3):                       # 1 warm-up, best of 2
            t0 = time.perf_counter()
            O.generator_forward(sd, lr, SCALE)
            dt = time.perf_counter() - t0
            if i and (best is None or dt < best):
                best = dt
    return {"metric": "x4 generator forward HR images/sec (216^2)", "value": batch / (ms * 1e-3), "unit": "HR images/s", "ms": ms,
            "tflops": 69.19e9 * batch / (ms * 1e-3) / 1e12,
            "config": {"workload": "SRADSGAN x4 generator forward, batch %d, LR 54^2 -> HR 216^2, bf16, eager launches" % batch},
            "cpu_baseline": {"value": batch / best, "unit": "HR images/s", "cores": cores, "kind": "port", "seconds": best,
                             "sample": "the same forward (oracle port, fp32, no_grad), best of 2 after 1 warm-up"}}


def gpu_comparator_bench(batch):
    """what a user of the reference has on this box TODAY — the same ATen program (the oracle port of the
    reference's step) run on the B200 through stock PyTorch / cuDNN, fp32 (TF32 convolutions, torch's default) and bf16 autocast
    + channels_last.  A comparator only: nothing of this build is on that path, and nothing of it is on this build's path."""
    import numpy as np
    import torch
    from oracle import sradsgan_oracle as O
    res = {}
    for mode in ("fp32", "bf16_autocast_channels_last"):
        try:
            dev = lambda sd: {k: v.cuda() for k, v in sd.items()}
            G = O.tie_upsampling(dev(O.make_state(O.generator_spec(SCALE), seed=0, init="ref")))
            st = O.TrainState(G, dev(O.make_state(O.discriminator_spec(), seed=1, init="ref")), dev(O.make_state(O.vgg_spec(), seed=2, init="fan")), SCALE)
            lr, hr = O.synthetic_batch(batch, SCALE, HR, seed=1234)
            lr, hr = lr.cuda(), hr.cuda()
            if mode != "fp32":
                lr, hr = lr.contiguous(memory_format=torch.channels_last), hr.contiguous(memory_format=torch.channels_last)
            rs = np.random.RandomState(1234)

            def one():
                alpha = torch.from_numpy(rs.random((batch, 1, 1, 1))).float().cuda()
                if mode == "fp32":
                    return O.train_step(st, lr, hr, alpha)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return O.train_step(st, lr, hr, alpha)
            for _ in range(2):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            res[mode] = {"value": batch / (ms * 1e-3), "unit": "HR images/s", "ms_per_step": ms}
            del st, G
        except Exception as e:
            res[mode] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
        torch.cuda.empty_cache()
    res["what"] = "oracle port of the reference's step on the same B200 through stock ATen/cuDNN kernels (torch %s), eager, 3 steps after 2 warm-up" % torch.__version__
    return res


def inference_bench(size, scale=9, tile=128, overlap=16, tile_batch=32):
    """Second half of BASELINE.json's metric: x9 generator inference on a synthetic size^2 LR tile (configs[3]) through
    overlapped tiling (the reference cannot run this: SGAM materialises an (HW)^2 attention).  Input resident on
    the device; output Mpix/s = (size*scale)^2 / time of one full pass (one warm-up pass over 2 tile batches)."""
    import torch
    from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup
    from sradsgan_b200.model.trainer import tiled_forward
    from sradsgan_b200.utils import weights_init_normal
    torch.manual_seed(0)
    G = GeneratorResNet(ResGroup, n_residual_blocks=12, n_basic_blocks=3, upscale_factor=scale)
    G.apply(weights_init_normal)
    G.cuda().eval()
    lr = torch.rand(1, 3, size, size, device="cuda")
    tiled_forward(G, lr[:, :, :2 * tile, :4 * tile].contiguous(), scale, tile, overlap, tile_batch)      # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = tiled_forward(G, lr, scale, tile, overlap, tile_batch)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    mpix = (size * scale) ** 2 / 1e6
    ok = bool(torch.isfinite(out[:, :, ::97, ::89]).all().item())
    del out
    torch.cuda.empty_cache()
    return {"metric": "x%d inference output Mpix/s" % scale, "value": mpix / sec, "unit": "Mpix/s", "seconds": sec,
            "config": {"workload": "SRADSGAN x%d generator, synthetic %dx%d LR -> %dx%d, overlapped tiles %d^2 (overlap %d LR px, "
                                   "feathered), %d tiles per forward, bf16" % (scale, size, size, size * scale, size * scale, tile, overlap, tile_batch)},
            "output_finite": ok}


def edsr_bench(batch, steps, warmup, peak_tf):
    """BASELINE.json configs[4]: EDSR(256 filters, 32 residual blocks) x4 L1 training on the same convolution kernels,
    batch 16, HR 216^2 / LR 54^2, bf16, CUDA-graph replay; inputs resident.  293.1 GFLOP per image forward,
    x3 for forward + input gradients + weight gradients."""
    import torch
    from sradsgan_b200.model.edsr import EDSR
    net = EDSR(trainer_args(model_name="EDSR", batch_size=batch, lr=1e-4, seed=0))
    net.build(init=True)
    g = torch.Generator().manual_seed(4321)
    hr = torch.rand(batch, 3, HR, HR, generator=g).cuda()
    lr = torch.nn.functional.interpolate(hr, size=HR // SCALE, mode="bicubic", align_corners=False).clamp(0, 1)
    for _ in range(max(warmup, 3)):
        out = net.graphed_step(lr, hr)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = net.graphed_step(lr, hr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flop_img = 3 * 293.1e9
    tf = flop_img * batch / (ms * 1e-3) / 1e12
    res = {"metric": "EDSR x4 train HR images/sec (216^2)", "value": batch / (ms * 1e-3), "unit": "HR images/s", "ms_per_step": ms,
           "tflops": tf, "frac_of_tensor_peak": tf / peak_tf, "loss_finite": bool(torch.isfinite(out["loss_G"]).item()),
           "gpu_launches_per_step": net._graph["launches"],
           "config": {"workload": "EDSR(256,32) x4 L1 training step, batch %d, HR 216^2 / LR 54^2, bf16, random init" % batch,
                      "flop_per_image": flop_img}}
    del net
    torch.cuda.empty_cache()
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 2)), (1 if args.warmup else 0)          # ~8 s per step on 16 cores: the run ends within a minute
    cb, mean = cpu_reference_step(batch=args.batch, steps=steps, warmup=warm)
    line = {"impl": "reference", "metric": "x4 train HR images/sec (216^2)", "value": cb["value"], "unit": "HR images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "SRADSGAN x4 full GAN training step (G+D, L1+adv+VGG19 perceptual, WGAN-GP), batch %d per GPU, "
                                                        "HR 216^2 / LR 54^2" % args.batch, "global_batch": args.batch * args.gpus,
                                            "parallelism": "dp%d" % args.gpus,
                                            "note": "the reference's own CPU path (%s) on the host cores, each step a bounded sample of the workload: "
                                                    "ONE rank's batch of %d; the reference is single-device, so its rate does not grow with --gpus"
                                                    % (cb["kind"], args.batch)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "HR images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="per-GPU batch (BASELINE config: 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eagerly launched step (train() with --no_graphs) instead of the graph replay")
    ap.add_argument("--no-inference", action="store_true", help="skip the x9 tiled-inference measurement (second half of the metric)")
    ap.add_argument("--no-edsr", action="store_true", help="skip the EDSR(256,32) x4 training measurement (BASELINE configs[4])")
    ap.add_argument("--no-comparator", action="store_true", help="skip the stock-PyTorch/cuDNN-on-this-GPU comparator and the G-forward line (configs[0])")
    ap.add_argument("--infer-size", type=int, default=2048, help="side of the synthetic LR image of the inference measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from sradsgan_b200 import _lib
    from sradsgan_b200.model.sradsgan import SRADSGAN
    be = _lib.backend()
    be.device_check()
    W = max(args.warmup, 3)
    K = args.steps
    B = args.batch

    net = SRADSGAN(trainer_args(batch_size=B, seed=0, graphs=not args.no_graph))
    net.build(init=True)
    g = torch.Generator().manual_seed(1234 + rank)
    hr_host = torch.rand(B, 3, HR, HR, generator=g).pin_memory()
    lr_host = torch.nn.functional.interpolate(hr_host, size=HR // SCALE, mode="bicubic", align_corners=False).clamp(0, 1).pin_memory()
    hr_dev = torch.empty_like(hr_host, device="cuda")
    lr_dev = torch.empty_like(lr_host, device="cuda")
    hr_dev.copy_(hr_host); lr_dev.copy_(lr_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # what one iteration of `net.train()` (main_sradsgan.py -> SRADSGAN.train -> _fit) executes: the CUDA-graph replay by
    # default, the eagerly launched step with --no-graph
    use_graph = not args.no_graph
    step_fn = net.step_fn()
    assert (step_fn == net.graphed_step) == use_graph

    # ---- per-kernel attribution: one eagerly-launched step with CUDA events around every library launch ----
    # (weight gradients normally run on a side stream next to the main one; for the attribution every kernel runs alone)
    from sradsgan_b200 import ops as _ops
    async_prev, _ops._WgradStream.enabled = _ops._WgradStream.enabled, False
    net.train_step(lr_dev, hr_dev)
    torch.cuda.synchronize()
    be.prof = []
    net.train_step(lr_dev, hr_dev)
    torch.cuda.synchronize()
    prof, be.prof = be.prof, None
    _ops._WgradStream.enabled = async_prev

    def timed(fn, lr, hr, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn(lr, hr)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), out

    # ---- device-resident timing (`value`) ----
    for _ in range(W):
        step_fn(lr_dev, hr_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = be.launch_count()
    ms, out = timed(step_fn, lr_dev, hr_dev, K)
    launches = (net._graph["launches"] * K) if use_graph else (be.launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None
    loss_ok = bool(torch.isfinite(out["loss_G"]).item() and torch.isfinite(out["loss_D"]).item())

    # ---- end-to-end: the loop body of train() with HOST batches (`e2e`): pinned host -> device copies of LR/HR inside the
    # step call, both losses read back every step ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        if use_graph:
            out = step_fn(lr_host, hr_host)                      # graphed_step stages its inputs from the loader's pinned batch
        else:
            lr_dev.copy_(lr_host, non_blocking=True); hr_dev.copy_(hr_host, non_blocking=True)     # reference :821-823
            out = step_fn(lr_dev, hr_dev)
        _ = (out["loss_G"].item(), out["loss_D"].item())
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = t.item()

    # ---- the other launch mode, beside it (N = 1 only: a short run) ----
    other = None
    if world == 1:
        other_fn = net.train_step if use_graph else net.graphed_step
        for _ in range(2):
            other_fn(lr_dev, hr_dev)
        oms, _ = timed(other_fn, lr_dev, hr_dev, 4)
        other = {"mode": "eager launches (train() --no_graphs)" if use_graph else "CUDA-graph replay (train() default)",
                 "ms_per_step": oms / 4, "value": B * 4 / (oms * 1e-3), "unit": "HR images/s"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-launch CUDA events recorded inside the timed region) ----
    agg = {}
    for rec in prof:
        kind, flops, nbytes, a, b = rec
        d = agg.setdefault(kind, [0.0, 0.0, 0, 0.0])
        d[0] += flops; d[1] += a.elapsed_time(b) * 1e-3; d[2] += 1; d[3] += nbytes
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained / hbm_gbs (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained, 6.65 TB/s (of fallback)"
    ms_step = ms / K
    kernels = {}
    for k, v in agg.items():
        kernels[k] = {"launches_per_step": v[2], "ms_per_step": v[1] * 1e3, "share_of_step": v[1] * 1e3 / ms_step}
        if v[0] > 0:
            kernels[k]["tflops"] = v[0] / v[1] / 1e12 if v[1] > 0 else None
        if v[3] > 0:
            kernels[k]["gbs"] = v[3] / v[1] / 1e9 if v[1] > 0 else None
            kernels[k]["frac_of_hbm"] = kernels[k]["gbs"] / peak_gbs if v[1] > 0 else None
    tensor_cls = {k: v for k, v in agg.items() if v[0] > 0 and "tcgen05" in k}
    dom = max(tensor_cls.items(), key=lambda kv: kv[1][1])[0] if tensor_cls else None
    roof = None
    if dom:
        ach = agg[dom][0] / agg[dom][1] / 1e12
        traffic, traffic_note = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
            if dom in tr:
                traffic = tr[dom]["traffic"]
                traffic_note = "%s: %s; algorithmic %d B" % (tr[dom]["layer"], tr["unit"], tr[dom]["algorithmic_bytes"])
        except Exception:
            pass
        hbm = {k: {"achieved": v["gbs"], "peak": peak_gbs, "unit": "GB/s", "frac": v["frac_of_hbm"], "ms_per_step": v["ms_per_step"],
                   "launches": v["launches_per_step"]} for k, v in kernels.items() if v.get("gbs")}
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "launches": agg[dom][2],
                "measured": "CUDA events around each launch of this kernel class on the launching stream, one eagerly launched step",
                "step_frac_of_tensor_roofline": (FLOP_PER_IMG * B * K / (ms * 1e-3) / 1e12) / peak_tf,
                "memory_bound_families": hbm}

    # the sub-lines (second half of the metric, second workload, configs[0], comparators, CPU baseline) are single-GPU
    # measurements: N = 1 only.  (At N > 1 the other ranks have left by now: a trainer built here must not enter a collective.)
    gfwd = comparator = None
    if not args.no_comparator and world == 1:
        try:
            gfwd = g_forward_bench(B, net)
        except Exception as e:
            gfwd = {"error": "%s: %s" % (type(e).__name__, e)}
    del net
    torch.cuda.empty_cache()
    if not args.no_comparator and world == 1:
        comparator = gpu_comparator_bench(B)

    infer = None
    if not args.no_inference and world == 1:
        infer = inference_bench(args.infer_size)

    edsr = None
    if not args.no_edsr and world == 1:
        try:
            edsr = edsr_bench(B, 5, 3, peak_tf)
        except Exception as e:      # the second workload must never take the headline number down with it
            edsr = {"error": "%s: %s" % (type(e).__name__, e)}

    cb = None
    if not args.no_cpu_baseline and world == 1:
        cb, _ = cpu_reference_step(batch=B, steps=1, warmup=1)

    imgs = B * world * K
    line = {"metric": "x4 train HR images/sec (216^2)", "value": imgs / (ms * 1e-3), "unit": "HR images/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "SRADSGAN x4 full GAN training step (G+D, L1+adv+VGG19 perceptual, WGAN-GP), batch %d per GPU, "
                                   "HR 216^2 / LR 54^2" % B, "global_batch": B * world, "parallelism": "dp%d" % world,
                       "l2": "per-step working set (activations+gradients, >2 GB) exceeds the 126 MB L2; no explicit flush",
                       "weights": "random init of the exact architecture (G 11.07M, D 4.70M, VGG19[:12] seeded)",
                       "flop_per_image": FLOP_PER_IMG, "loss_finite": loss_ok, "cuda_graph": use_graph,
                       "entry_point": "the loop body of SRADSGAN.train() (net.step_fn())"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": imgs / e2e_s, "unit": "HR images/s", "h2d_bytes_per_step": int(lr_host.numel() * 4 + hr_host.numel() * 4),
                    "d2h_bytes_per_step": 8},
            "other_launch_mode": other,
            "roofline": roof, "kernels": kernels, "g_forward": gfwd, "gpu_comparator": comparator, "inference": infer, "edsr": edsr,
            "cpu_baseline": cb}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
