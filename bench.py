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
    """The reference's CPU path: oracle/sradsgan_oracle.train_step (same ATen ops as the reference modules,
    pinned to them by tests/golden) on all host cores; `batch` images of the SAME x4 216^2 workload per step."""
    import numpy as np
    import torch
    from oracle import sradsgan_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    G = O.tie_upsampling(O.make_state(O.generator_spec(SCALE), seed=0, init="ref"))
    D = O.make_state(O.discriminator_spec(), seed=1, init="ref")
    V = O.make_state(O.vgg_spec(), seed=2, init="fan")
    st = O.TrainState(G, D, V, SCALE)
    lr, hr = O.synthetic_batch(batch, SCALE, HR, seed=1234)
    rs = np.random.RandomState(1234)
    times = []
    for i in range(warmup + steps):
        alpha = torch.from_numpy(rs.random((batch, 1, 1, 1))).float()
        t0 = time.perf_counter()
        O.train_step(st, lr, hr, alpha)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return {"value": batch / mean, "unit": "HR images/s", "cores": cores, "kind": "port",
            "sample": "%d full G+D training step(s) of the same x4 216^2 workload at batch %d (of %d), fp32, %d threads; "
                      "%.2f s/step" % (steps, batch, BATCH, cores, mean)}, mean


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
    cb, mean = cpu_reference_step(batch=2, steps=max(1, min(args.steps, 3)), warmup=1 if args.warmup else 0)
    line = {"impl": "reference", "metric": "x4 train HR images/sec (216^2)", "value": cb["value"], "unit": "HR images/s",
            "n_gpus": args.gpus, "steps": max(1, min(args.steps, 3)), "warmup": 1 if args.warmup else 0,
            "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "SRADSGAN x4 full GAN training step (G+D, L1+adv+VGG19), HR 216^2 / LR 54^2, "
                                            "reference CPU path on a bounded sample (batch 2 per step)"},
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
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-inference", action="store_true", help="skip the x9 tiled-inference measurement (second half of the metric)")
    ap.add_argument("--no-edsr", action="store_true", help="skip the EDSR(256,32) x4 training measurement (BASELINE configs[4])")
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

    net = SRADSGAN(trainer_args(batch_size=B, seed=0))
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

    use_graph = not args.no_graph
    step_fn = net.graphed_step if use_graph else net.train_step

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

    # ---- device-resident timing (`value`) ----
    for _ in range(W):
        step_fn(lr_dev, hr_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = be.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        out = step_fn(lr_dev, hr_dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = (net._graph["launches"] * K) if use_graph else (be.launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    loss_ok = bool(torch.isfinite(out["loss_G"]).item() and torch.isfinite(out["loss_D"]).item())

    # ---- end-to-end through the public API with host buffers (`e2e`) ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        lr_dev.copy_(lr_host, non_blocking=True)
        hr_dev.copy_(hr_host, non_blocking=True)
        out = step_fn(lr_dev, hr_dev)
        _ = (out["loss_G"].item(), out["loss_D"].item())
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = t.item()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (per-launch CUDA events recorded inside the timed region) ----
    agg = {}
    for kind, flops, a, b in prof:
        d = agg.setdefault(kind, [0.0, 0.0, 0])
        d[0] += flops; d[1] += a.elapsed_time(b) * 1e-3; d[2] += 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    ms_step = ms / K
    kernels = {k: {"launches_per_step": v[2], "ms_per_step": v[1] * 1e3, "tflops": (v[0] / v[1] / 1e12) if v[1] > 0 else None,
                   "share_of_step": v[1] * 1e3 / ms_step} for k, v in agg.items()}
    dom = max(agg.items(), key=lambda kv: kv[1][1])[0] if agg else None
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
        roof = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "launches": agg[dom][2],
                "measured": "CUDA events around each launch of this kernel class on the launching stream, one eagerly launched step",
                "step_frac_of_tensor_roofline": (FLOP_PER_IMG * B * K / (ms * 1e-3) / 1e12) / peak_tf}

    # the sub-lines (second half of the metric, second workload, CPU baseline) are single-GPU measurements: N = 1 only.
    # (At N > 1 the other ranks have left by now: a trainer built here must not enter a collective.)
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
        cb, _ = cpu_reference_step(batch=2, steps=1, warmup=0)

    imgs = B * world * K
    line = {"metric": "x4 train HR images/sec (216^2)", "value": imgs / (ms * 1e-3), "unit": "HR images/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "SRADSGAN x4 full GAN training step (G+D, L1+adv+VGG19 perceptual, WGAN-GP), batch %d per GPU, "
                                   "HR 216^2 / LR 54^2" % B, "global_batch": B * world, "parallelism": "dp%d" % world,
                       "l2": "per-step working set (activations+gradients, >2 GB) exceeds the 126 MB L2; no explicit flush",
                       "weights": "random init of the exact architecture (G 11.07M, D 4.70M, VGG19[:12] seeded)",
                       "flop_per_image": FLOP_PER_IMG, "loss_finite": loss_ok, "cuda_graph": use_graph},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": imgs / e2e_s, "unit": "HR images/s", "h2d_bytes_per_step": int(lr_host.numel() * 4 + hr_host.numel() * 4),
                    "d2h_bytes_per_step": 8},
            "roofline": roof, "kernels": kernels, "inference": infer, "edsr": edsr, "cpu_baseline": cb}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
