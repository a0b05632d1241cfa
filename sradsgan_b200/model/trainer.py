"""Trainer with the reference's entry points — `SRADSGAN(args).train()`, `.gradient_penalty()`,
`.mfe_test_single()`, `.validate()`, `.mfeNew_validate()`, `.mfeNew_validateByClass()`, `save/load_*`
(reference model/sradsgan.py:510-1640) — driving the B200-native networks.

What is kept bit-for-bit in structure: loss_G = L1 + 1e-2*L1(VGG) + 1e-3*(-mean D(G(x)));
loss_D = -mean D(hr) + mean D(G(x).detach()) + lambda*GP with the GP ALSO back-propagated once un-weighted
(effective weight 1+lambda, reference :639 + :884-886); per-pixel GP norm over the 3 colour channels (:630);
D always in train mode (4 BatchNorm stat updates per step); Adam(2e-4, .9, .999); clamp of every D
parameter to +-clip_value (:891-892).
What is deliberately not reproduced because it changes no result: D / VGG weight gradients during the G
step (zeroed / never used, :865), the second traversal of the GP double-backward graph (the two
backward passes are fused into one with weight 1+lambda), per-iteration `.item()` syncs and empty_cache().
"""
import os
import time
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, dp, ops
from ..optim import FlatAdam
from ..utils import CsvLogger, eval_metrics_u8, psnr, save_img1, weights_init_normal


class SRADSGAN(object):
    def __init__(self, args):
        # parameters (reference :513-556)
        for k in ("model_name", "train_dataset", "test_dataset", "crop_size", "test_crop_size", "hr_height", "hr_width",
                  "num_threads", "num_channels", "scale_factor", "epoch", "num_epochs", "save_epochs", "batch_size",
                  "test_batch_size", "lr", "b1", "b2", "data_dir", "root_dir", "save_dir", "gpu_mode", "n_cpu",
                  "sample_interval", "clip_value", "lambda_gp", "gp", "penalty_type", "grad_penalty_Lp_norm",
                  "loss_Lp_norm", "weight_gan", "weight_content", "max_train_samples"):
            setattr(self, k, getattr(args, k))
        self.relative = args.relativeGan
        # new flags (defaults keep the reference behaviour)
        self.precision = getattr(args, "precision", "bf16")
        self.synthetic_steps = getattr(args, "synthetic_steps", 0)
        self.log_interval = getattr(args, "log_interval", 50)
        self.vgg_state = getattr(args, "vgg_state", None)
        self.seed = getattr(args, "seed", 0)
        self.gpu_input_pipeline = getattr(args, "gpu_input_pipeline", False)
        # train() replays the captured CUDA graph of the iteration (graphed_step) unless --no_graphs: the drop-in entry point
        # owns the fast path; eager launches remain for debugging / profiling
        self.use_graphs = bool(getattr(args, "graphs", True))
        # chain training (the paper's x2 -> x3 -> x4 ... warm starts; reference :716-721 does it by hand-editing two paths)
        self.pretrained_G = getattr(args, "pretrained_G", None)
        self.pretrained_D = getattr(args, "pretrained_D", None)
        ops.set_precision(self.precision)
        if self.relative:
            raise NotImplementedError("relativistic GAN branch (reference :840-844) is unreachable from the CLI and not built")
        self.cuda = torch.cuda.is_available()
        if not self.cuda and _lib.backend().name == "cuda":
            raise RuntimeError("sradsgan_b200 needs a CUDA (sm_100) device; there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if self.cuda else torch.device("cpu")
        self.log_dict = OrderedDict()
        self.logger = None
        self.generator = self.discriminator = self.feature_extractor = None
        self.optimizer_G = self.optimizer_D = None
        self._alpha_override = None
        self._graph = None
        self._pack_plans = None          # ops.PackPlan per network, built by _capture (batched weight re-packing)

    # ------------------------------------------------------------------------------------------
    # construction
    # ------------------------------------------------------------------------------------------
    def new_generator(self):
        from .sradsgan import GeneratorResNet, ResGroup
        return GeneratorResNet(ResGroup, n_residual_blocks=12, n_basic_blocks=3, rla_mode='CA-SA', bla_mode='CA-SA',
                               ga_mode='CA-SA', pool_mode='Avg|Max', upscale_factor=self.scale_factor)   # :669-671

    def build(self, init=True):
        """networks, losses and optimisers (reference :669-725)"""
        from .sradsgan import Discriminator, FeatureExtractor, GANLoss
        torch.manual_seed(self.seed)
        self.generator = self.new_generator()
        self.discriminator = Discriminator()
        vsd = torch.load(self.vgg_state, map_location="cpu") if isinstance(self.vgg_state, str) else self.vgg_state
        self.feature_extractor = FeatureExtractor(state_dict=vsd)
        self.criterion_raGAN = GANLoss(gan_type='wgan-gp', real_label_val=1.0, fake_label_val=0.0)
        if init and self.epoch == 0:
            self.generator.apply(weights_init_normal)        # :713-714
            self.discriminator.apply(weights_init_normal)
        for m in (self.generator, self.discriminator, self.feature_extractor):
            m.to(self.device)
        self._make_optimizers()

    def _make_optimizers(self):
        self.optimizer_G = FlatAdam(self.generator, lr=self.lr, betas=(self.b1, self.b2), chunk_of=dp.generator_chunk_key)
        self.optimizer_D = FlatAdam(self.discriminator, lr=self.lr, betas=(self.b1, self.b2),
                                    clamp=(-self.clip_value, self.clip_value))
        dp.broadcast_parameters(self.optimizer_G)
        dp.broadcast_parameters(self.optimizer_D)
        # The G bucket (44 MB) can be all-reduced chunk by chunk (one chunk per ResGroup) while backward is still
        # running (SR_DP_OVERLAP=1, eager launches only); at ~0.1 ms per bucket over NVLink the default is one
        # all-reduce per network after its backward.
        self.reducer_G = dp.BucketReducer(self.optimizer_G, overlap=os.environ.get("SR_DP_OVERLAP", "0") == "1")
        self.reducer_D = dp.BucketReducer(self.optimizer_D, overlap=False)

    def criterion_content(self, a, b):
        return ops.diff_mean_loss(a, b, 1 if self.loss_Lp_norm == "L1" else 2)     # :685-688 (nn.L1Loss / nn.MSELoss)

    # ------------------------------------------------------------------------------------------
    # WGAN-GP (reference :595-641)
    # ------------------------------------------------------------------------------------------
    def _gp_graph(self, discriminator, real_samples, fake_samples, grad_penalty_Lp_norm, penalty_type):
        b = real_samples.size(0)
        if self._alpha_override is not None:
            alpha = self._alpha_override.to(real_samples.device, torch.float32).view(b, 1, 1, 1)
        else:
            alpha = torch.from_numpy(np.random.random((b, 1, 1, 1))).float().to(real_samples.device)   # :609
        interpolates = ops.gp_interpolates(real_samples, fake_samples, alpha).requires_grad_(True)     # :611
        d_interpolates = discriminator(interpolates)
        grad_outputs = torch.ones_like(d_interpolates)
        with ops.input_grad_only():      # only d/d(interpolates) is wanted here; weight gradients come from .backward()
            gradients = torch.autograd.grad(outputs=d_interpolates, inputs=interpolates, grad_outputs=grad_outputs,
                                            create_graph=True, retain_graph=True, only_inputs=True)[0]     # :621
        # per-pixel norm over the colour channels (:623-630), 'LS' (n-1)^2 or hinge (:632-635), mean (:637): one reduction
        # kernel forward, one elementwise kernel backward (its output is the cotangent of the double backward through D)
        return ops.gp_penalty(gradients, grad_penalty_Lp_norm, penalty_type)

    def gradient_penalty(self, discriminator, real_samples, fake_samples, grad_penalty_Lp_norm='L2', penalty_type='LS'):
        """Same contract as the reference: back-propagates the un-weighted penalty itself and returns it."""
        gp = self._gp_graph(discriminator, real_samples, fake_samples, grad_penalty_Lp_norm, penalty_type)
        gp.backward(retain_graph=True)                                                                 # :639
        return gp

    # ------------------------------------------------------------------------------------------
    # one iteration (reference :829-892)
    # ------------------------------------------------------------------------------------------
    def _g_phase(self, imgs_lr, imgs_hr):
        """generator forward + loss + backward (reference :829-857): leaves the gradients in optimizer_G.flat_grad"""
        G, D, Fx = self.generator, self.discriminator, self.feature_extractor
        mark = getattr(self, "_phase_mark", None) or (lambda name: None)
        mark("start")
        self._repack()
        # the loader's NCHW fp32 batches -> NHWC fp32 once per iteration (one layout kernel each); VGG, the critic, the L1 loss
        # and the WGAN-GP interpolation all read these
        imgs_lr, imgs_hr = ops.to_compute(imgs_lr), ops.to_compute(imgs_hr)
        self.optimizer_G.zero_grad()
        for p in self.optimizer_D.params:
            p.requires_grad_(False)          # skip D's weight gradients in the G step (discarded by the reference, :865)
        # VGG features of the HR batch (:837, detached) are independent of the generator; SR_VGG_ASYNC=1 computes them on the
        # side stream during the generator's forward pass.  Measured SLOWER (28.6 vs 27.9 ms per step: the large 216^2 VGG
        # kernels take SMs from the generator's critical path), so the default keeps them on the main stream.
        overlap_vgg = imgs_hr.is_cuda and os.environ.get("SR_VGG_ASYNC", "0") == "1"
        if overlap_vgg:
            side = ops.side_stream(imgs_hr.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():
                real_features = Fx(imgs_hr)
        gen_hr = G(imgs_lr)                                                         # :832
        mark("G_fwd")
        pixel_loss_G = self.criterion_content(gen_hr, imgs_hr)                      # :834
        gen_features = Fx(gen_hr)                                                   # :836
        if overlap_vgg:
            torch.cuda.current_stream().wait_stream(side)
            real_features.record_stream(torch.cuda.current_stream())
        else:
            with torch.no_grad():
                real_features = Fx(imgs_hr)                                         # :837
        loss_content = self.criterion_content(gen_features, real_features)         # :838
        mark("VGG_fwd_x2")
        loss_gan = self.criterion_raGAN(D(gen_hr), True)                            # :847-848
        loss_G = pixel_loss_G + self.weight_content * loss_content + self.weight_gan * loss_gan   # :852
        mark("D_adv_fwd")
        self.reducer_G.arm()
        loss_G.backward()
        ops.wgrad_join()
        mark("G_step_backward")
        for p in self.optimizer_D.params:
            p.requires_grad_(True)
        return {"loss_G": loss_G.detach(), "pixel": pixel_loss_G.detach(), "content": loss_content.detach(),
                "adv": loss_gan.detach(), "gen_hr": gen_hr.detach(), "_hr_nhwc": imgs_hr}

    def _repack(self):
        """packed bf16 operands of every G and D convolution weight, refreshed from the fp32 masters by one launch per
        network at the start of the step (ops.PackPlan; built by _capture after its eager warm-up steps)"""
        plans = getattr(self, "_pack_plans", None)
        if not plans:
            return
        if not all(pl.valid() for pl in plans):
            self._pack_plans = None
            return
        for pl in plans:
            pl.repack()

    def _d_phase(self, imgs_hr, gen_det, fuse_gp_backward=True):
        """discriminator losses + WGAN-GP + backward (reference :865-886): gradients in optimizer_D.flat_grad"""
        D = self.discriminator
        mark = getattr(self, "_phase_mark", None) or (lambda name: None)
        self.optimizer_D.zero_grad()                                                # :865
        loss_real = self.criterion_raGAN(D(imgs_hr), True)                          # :876
        loss_fake = self.criterion_raGAN(D(gen_det), False)                         # :877
        loss_D = loss_real + loss_fake
        mark("D_real_fake_fwd")
        if self.gp:
            if fuse_gp_backward:
                gp = self._gp_graph(D, imgs_hr, gen_det, self.grad_penalty_Lp_norm, self.penalty_type)
                mark("GP_fwd_and_grad")
                (loss_D + (1.0 + self.lambda_gp) * gp).backward()                   # == :639 followed by :886
                mark("D_step_backward")
                loss_D = loss_D.detach() + self.lambda_gp * gp.detach()
            else:
                gp = self.gradient_penalty(D, imgs_hr, gen_det, self.grad_penalty_Lp_norm, self.penalty_type)
                loss_D = loss_D + self.lambda_gp * gp                               # :884
                loss_D.backward()                                                   # :886
        else:
            gp = torch.zeros((), device=imgs_hr.device)
            loss_D.backward()
        ops.wgrad_join()
        return {"loss_D": loss_D.detach(), "gp": gp.detach()}

    def train_step(self, imgs_lr, imgs_hr, fuse_gp_backward=True):
        mark = getattr(self, "_phase_mark", None) or (lambda name: None)
        out = self._g_phase(imgs_lr, imgs_hr)
        scale = self.reducer_G.finish()
        self.optimizer_G.step(grad_scale=scale)                                     # :857-858
        mark("adam_G")
        out.update(self._d_phase(out["_hr_nhwc"], out["gen_hr"], fuse_gp_backward))
        self.reducer_D.arm()
        scale = self.reducer_D.finish()
        self.optimizer_D.step(grad_scale=scale)                                     # :887 + clamp :891-892 (fused)
        mark("adam_D")
        return out

    # ------------------------------------------------------------------------------------------
    # CUDA-graph replay of the whole iteration
    # ------------------------------------------------------------------------------------------
    def graphed_step(self, imgs_lr, imgs_hr):
        """train_step captured ONCE into CUDA graphs (~6k kernel launches -> a few cudaGraphLaunch) and replayed;
        inputs / the GP interpolation factors are staged into static device buffers, the Adam step counters live
        on the device.  One process: a single graph.  Data parallel: four graph segments (G phase | Adam_G | D phase |
        Adam_D) with the two NCCL gradient all-reduces issued BETWEEN the replays (collectives are never captured) and the
        generator's all-reduce + Adam running next to the D phase (_replay_dp).  Re-capture happens when the input shape or
        a learning rate changes."""
        key = (tuple(imgs_lr.shape), tuple(imgs_hr.shape), self.optimizer_G.param_groups[0]["lr"], self.optimizer_D.param_groups[0]["lr"])
        if self._graph is None or self._graph["key"] != key:
            self._capture(imgs_lr, imgs_hr, key)
        g = self._graph
        g["lr"].copy_(imgs_lr, non_blocking=True)
        g["hr"].copy_(imgs_hr, non_blocking=True)
        # GP interpolation factors (reference :609, numpy RNG): two pinned staging buffers used alternately, each guarded by
        # an event recorded after its host->device copy — the host runs many replays ahead of the device and must not
        # overwrite a buffer whose copy is still queued
        slot = g["alpha_slot"] = 1 - g.get("alpha_slot", 1)
        g["alpha_evt"][slot].synchronize()
        g["alpha_host"][slot].copy_(torch.from_numpy(np.random.random((imgs_hr.size(0), 1, 1, 1))).float())
        g["alpha"].copy_(g["alpha_host"][slot], non_blocking=True)
        g["alpha_evt"][slot].record()
        if len(g["graphs"]) == 1:
            g["graphs"][0].replay()
        else:
            self._replay_dp(g)
        self.optimizer_G.step_count += 1
        self.optimizer_D.step_count += 1
        # the replay updated the fp32 masters through raw pointers: eager users of cached packed operands (validate(),
        # generator(x)) must re-pack (the graph itself re-packs its own operands at the start of every replay)
        self.optimizer_G.touch()
        self.optimizer_D.touch()
        return g["out"]

    def _replay_dp(self, g):
        """data parallel: G phase | all-reduce(G) + Adam_G on the communication stream NEXT TO the D phase (which reads neither
        the generator's gradients nor its updated weights: gen_hr is detached, reference :857-861) | all-reduce(D) | Adam_D.
        Collectives are never captured; the main stream joins the communication stream before the next G phase.
        SR_DP_OVERLAP=0 restores the serial order (G phase | all-reduce | Adam_G + D phase | all-reduce | Adam_D)."""
        g_phase, adam_g, d_phase, adam_d = g["graphs"]
        main = torch.cuda.current_stream()
        comm = g.get("comm_stream")
        g_phase.replay()
        if comm is not None:
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                dp.all_reduce_flat(self.optimizer_G.flat_grad)
                adam_g.replay()
        else:
            dp.all_reduce_flat(self.optimizer_G.flat_grad)
            adam_g.replay()
        d_phase.replay()
        dp.all_reduce_flat(self.optimizer_D.flat_grad)
        adam_d.replay()
        if comm is not None:
            main.wait_stream(comm)

    def _capture(self, imgs_lr, imgs_hr, key):
        dev = self.device                 # the batch may still be in (pinned) host memory: the static inputs live on the device
        world = dp.world_size()
        st = {"key": key, "lr": torch.empty(imgs_lr.shape, dtype=torch.float32, device=dev),
              "hr": torch.empty(imgs_hr.shape, dtype=torch.float32, device=dev),
              "alpha": torch.empty(imgs_hr.size(0), 1, 1, 1, device=dev),
              "alpha_host": [torch.empty(imgs_hr.size(0), 1, 1, 1).pin_memory() for _ in range(2)],
              "alpha_evt": [torch.cuda.Event(), torch.cuda.Event()]}
        st["lr"].copy_(imgs_lr); st["hr"].copy_(imgs_hr); st["alpha"].uniform_()
        prev_override = self._alpha_override
        self._alpha_override = st["alpha"]
        # snapshot everything a step mutates, so that warm-up + capture do not advance the training state
        snap = [t.clone() for t in self._mutable_state()]
        red = (self.reducer_G, self.reducer_D)
        self.reducer_G, self.reducer_D = dp.NullReducer(world), dp.NullReducer(world)     # no collectives while warming up / capturing
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.train_step(st["lr"], st["hr"])
            torch.cuda.current_stream().wait_stream(s)
            if os.environ.get("SR_PACK_PLAN", "1") == "1":
                self._pack_plans = [ops.PackPlan(self.optimizer_G.params), ops.PackPlan(self.optimizer_D.params)]
            n0 = _lib.backend().launch_count()
            # SR_GRAPH_PRIORITY=1: capture the main chain on a HIGH-priority stream, so that its kernel nodes win the block scheduler
            # against the weight-gradient branches forked onto the (default-priority) side streams
            cap = {"stream": torch.cuda.Stream(priority=-1)} if os.environ.get("SR_GRAPH_PRIORITY", "0") == "1" else {}
            if world == 1 and os.environ.get("SR_DP_FORCE_SEGMENTS", "0") != "1":      # (the knob: scripts/dp_overlap_check.py on one GPU)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local", **cap):
                    st["out"] = self.train_step(st["lr"], st["hr"])
                graphs = [graph]
            else:
                g1, ga, g2, g3 = (torch.cuda.CUDAGraph() for _ in range(4))
                with torch.cuda.graph(g1, capture_error_mode="thread_local"):
                    out = self._g_phase(st["lr"], st["hr"])
                with torch.cuda.graph(ga, pool=g1.pool(), capture_error_mode="thread_local"):
                    self.optimizer_G.step(grad_scale=1.0 / world)
                with torch.cuda.graph(g2, pool=g1.pool(), capture_error_mode="thread_local"):
                    out.update(self._d_phase(out["_hr_nhwc"], out["gen_hr"]))
                with torch.cuda.graph(g3, pool=g1.pool(), capture_error_mode="thread_local"):
                    self.optimizer_D.step(grad_scale=1.0 / world)
                st["out"] = out
                graphs = [g1, ga, g2, g3]
                st["comm_stream"] = torch.cuda.Stream() if os.environ.get("SR_DP_OVERLAP", "1") == "1" else None
            st["launches"] = _lib.backend().launch_count() - n0     # library kernels per replay
            torch.cuda.synchronize()
        finally:
            self.reducer_G, self.reducer_D = red
        for t, c in zip(self._mutable_state(), snap):
            t.copy_(c)
        self.optimizer_G.step_count -= 3
        self.optimizer_D.step_count -= 3
        ops.bump_weight_generation()
        self._alpha_override = prev_override if prev_override is not st["alpha"] else None
        self._alpha_static = st["alpha"]
        st["graphs"] = graphs
        st["pack_plans"] = getattr(self, "_pack_plans", None)     # the captured launches read the plans' device tables
        self._graph = st

    def _mutable_state(self):
        oG, oD = self.optimizer_G, self.optimizer_D
        bufs = [b for b in self.discriminator.buffers()]
        return [oG.flat_param, oG.exp_avg, oG.exp_avg_sq, oG.step_t, oD.flat_param, oD.exp_avg, oD.exp_avg_sq, oD.step_t] + bufs

    # ------------------------------------------------------------------------------------------
    # data
    # ------------------------------------------------------------------------------------------
    def load_dataset(self, dataset='train', max_samples=20000):
        """reference :643-656. Folder datasets are outside the hot path: a synthetic source
        is used when `synthetic_steps` > 0, otherwise a minimal PIL folder reader.  Data parallel: the TRAINING set is
        sharded per rank (DistributedSampler; `_train_sampler.set_epoch` is called by the loop) — every rank sees a disjoint
        slice per epoch, so the effective batch is batch_size * world.  Evaluation sets are read in full by every rank."""
        from ..data import DevicePrefetcher, FolderHRDataset, FolderSRDataset, SyntheticSRDataset
        train = dataset == 'train'
        bs = self.batch_size if train else self.test_batch_size
        device_pipeline = getattr(self, "gpu_input_pipeline", False) and not self.synthetic_steps
        if self.synthetic_steps:
            ds = SyntheticSRDataset(self.synthetic_steps * bs, self.crop_size, self.scale_factor, seed=1234 + _rank())
        else:
            names = self.train_dataset if train else self.test_dataset
            cls = FolderHRDataset if device_pipeline else FolderSRDataset
            ds = cls(self.data_dir, names, self.crop_size, self.scale_factor, max_samples=max_samples)
        sampler = None
        shuffle = train and not self.synthetic_steps
        if train and dp.is_dist() and not self.synthetic_steps:
            from torch.utils.data.distributed import DistributedSampler
            sampler = DistributedSampler(ds, num_replicas=dp.world_size(), rank=_rank(), shuffle=True, seed=self.seed, drop_last=True)
            shuffle = False
        if train:
            self._train_sampler = sampler
        gen = torch.Generator().manual_seed(self.seed * 7919 + _rank())      # shuffle order / worker seeds differ per rank
        loader = torch.utils.data.DataLoader(ds, num_workers=0 if self.synthetic_steps else self.num_threads, batch_size=bs,
                                             shuffle=shuffle, sampler=sampler, drop_last=True, pin_memory=self.cuda, generator=gen)
        # --gpu_input_pipeline: workers only decode + crop; LR / bicubic synthesis (PIL-exact) and the copies run on the device,
        # double-buffered behind the training step
        return DevicePrefetcher(loader, self.device, self.scale_factor) if device_pipeline else loader

    # ------------------------------------------------------------------------------------------
    # training loop (reference :658-1056)
    # ------------------------------------------------------------------------------------------
    def step_fn(self):
        """what one iteration of train() executes: the CUDA-graph replay (default) or the eagerly launched step (--no_graphs)"""
        return self.graphed_step if (self.use_graphs and self.cuda) else self.train_step

    def train(self):
        self.build()
        model_dir = os.path.join(self.save_dir, 'model')
        os.makedirs(model_dir, exist_ok=True)
        if self.epoch != 0:                                                         # :705-710
            self.load_epoch_network(model_dir + '/generator_param_epoch_%d.pkl' % self.epoch, self.generator, strict=True)
            self.load_epoch_network(model_dir + '/discriminator_param_epoch_%d.pkl' % self.epoch, self.discriminator, strict=True)
        elif self.pretrained_G or self.pretrained_D:                                # :716-721 (chain training warm start)
            self.load_pretrained(self.pretrained_G, self.pretrained_D)
        avg_loss_G, avg_loss_D = self._fit(model_dir, {"generator": self.generator, "discriminator": self.discriminator})
        if _rank() == 0:
            self.save_model(epoch=None)                                              # :1056
        return avg_loss_G, avg_loss_D

    def _fit(self, model_dir, nets):
        """the epoch loop of the reference's train() (:804-1036), shared with the EDSR trainer (model/edsr.py:207-372).
        `nets`: label -> module saved every epoch; `nets['generator']` is the one the rollback heuristic reloads."""
        self.logger = CsvLogger(os.path.join(self.save_dir, 'logs')) if _rank() == 0 else None
        lr_sz = self.crop_size // self.scale_factor
        run = self.step_fn()
        graphs = run == self.graphed_step
        if not graphs:       # the reference's pre-allocated staging tensors (:739-741); graphed_step owns its static buffers
            input_lr = torch.empty(self.batch_size, self.num_channels, lr_sz, lr_sz, device=self.device)
            input_hr = torch.empty(self.batch_size, self.num_channels, self.crop_size, self.crop_size, device=self.device)
        dataloader = self.load_dataset('train', max_samples=self.max_train_samples)
        print('Training is started.')
        step, start_time = 0, time.time()
        epoch = self.epoch
        # reference :795-800 — best-so-far of the four validation metrics and the no-improvement counter
        best = {"psnr": 0.0, "ssim": 0.0, "ergas": 10000.0, "lpips": 10000.0, "step": 0, "no_improve": 0}
        has_D = self.optimizer_D is not None
        avg_loss_G, avg_loss_D = [], []
        while epoch < self.num_epochs and self.lr >= 0.00001:                       # :804
            if getattr(self, "_train_sampler", None) is not None:
                self._train_sampler.set_epoch(epoch)
            sum_G = torch.zeros((), device=self.device)
            sum_D = torch.zeros((), device=self.device)
            n_it = 0
            for i, batch in enumerate(dataloader):
                if graphs:                                                          # one copy: loader batch -> the graph's inputs
                    imgs_lr, imgs_hr = batch[0], batch[1]
                else:
                    imgs_lr = input_lr.copy_(batch[0], non_blocking=True)           # :821-823
                    imgs_hr = input_hr.copy_(batch[1], non_blocking=True)
                out = run(imgs_lr, imgs_hr)
                sum_G += out["loss_G"]; sum_D += out["loss_D"]; n_it += 1
                step += 1
                if self.logger is not None and (step % self.log_interval == 0 or step == 1):
                    lg, ld = out["loss_G"].item(), out["loss_D"].item()             # the only host sync, every N steps
                    print("[Epoch %d/%d] [Batch %d/%d] [D loss: %f] [G loss: %f]" % (epoch, self.num_epochs, i, len(dataloader), ld, lg))
                    self.logger.scalar_summary('loss_G', lg, step)
                    if has_D:
                        self.logger.scalar_summary('loss_D', ld, step)
                    if step % self.sample_interval == 0 or step == 1:
                        self.logger.print_format_results('train', OrderedDict(
                            model=self.model_name, epoch=epoch, iters=step, G_lr=self.optimizer_G.param_groups[0]['lr'],
                            D_lr=self.optimizer_D.param_groups[0]['lr'] if has_D else 0.0, time=time.time() - start_time,
                            G_loss=lg, D_loss=ld,
                            srwgan_psnr=psnr(out["gen_hr"][0].float().cpu(), batch[1][0].float().cpu())))
            avg_loss_G.append((sum_G / max(n_it, 1)).item())
            avg_loss_D.append((sum_D / max(n_it, 1)).item())
            val = self.validate(epoch=epoch, mode='train', save_img=((epoch + 1) % self.save_epochs == 0))
            self._update_best(best, val, epoch)
            if _rank() == 0:
                for label, net in nets.items():
                    self.save_epoch_network(model_dir, net, label, epoch + 1)        # :1005-1008
            if dp.is_dist():
                torch.distributed.barrier()          # a rollback below reads the checkpoint rank 0 has just written
            epoch += 1
            if best["no_improve"] >= 5:                                              # :1011-1036 LR halving heuristic
                self.load_epoch_network(model_dir + '/generator_param_epoch_%d.pkl' % (best["step"] + 1), nets["generator"])
                for g in self.optimizer_G.param_groups:
                    g["lr"] /= 2.0
                if has_D and self.lr < 0.0001:
                    for g in self.optimizer_D.param_groups:
                        g["lr"] /= 2.0
                self.lr /= 2.0
                epoch = best["step"] + 1
                best["no_improve"] = 0
        print("Training is finished.")
        return avg_loss_G, avg_loss_D

    @staticmethod
    def _update_best(best, val, epoch):
        """reference :986-1003: the counter is reset when PSNR, else SSIM, else ERGAS, else LPIPS improves (in that `elif`
        order); metrics this build does not compute are None and never participate.  A validation pass that produced no
        samples (synthetic training, no test folder) returns None and leaves the heuristic untouched — otherwise every
        epoch would count as 'no improvement' and the run would roll back to epoch 1 for ever."""
        if val is None:
            return
        v_psnr, v_ssim, v_ergas, v_lpips = val
        if v_psnr is not None and v_psnr - best["psnr"] > 0:
            best.update(psnr=v_psnr, no_improve=0, step=epoch)
        elif v_ssim is not None and v_ssim - best["ssim"] > 0:
            best.update(ssim=v_ssim, no_improve=0, step=epoch)
        elif v_ergas is not None and v_ergas - best["ergas"] < 0:
            best.update(ergas=v_ergas, no_improve=0, step=epoch)
        elif v_lpips is not None and v_lpips - best["lpips"] < 0:
            best.update(lpips=v_lpips, no_improve=0, step=epoch)
        else:
            best["no_improve"] += 1

    # ------------------------------------------------------------------------------------------
    # evaluation / inference (reference :1058-1194, :1259-1640)
    # ------------------------------------------------------------------------------------------
    def _eval_loader(self, names=None):
        prev = self.test_dataset
        try:
            if names is not None:
                self.test_dataset = names
            return self.load_dataset('test')
        except (FileNotFoundError, OSError):
            return None
        finally:
            self.test_dataset = prev

    @torch.no_grad()
    def _evaluate(self, loader, save_dir=None, epoch=0, classname=None):
        """generator over one evaluation loader; the per-image metrics of the reference's loops (:1111-1122, :1481-1494 —
        skimage compare_mse / compare_psnr / compare_ssim on the uint8 images, compare_ergas2) for the SR output AND the
        bicubic baseline stay on the device (utils.eval_metrics_u8); ONE read-back per loader.
        LPIPS needs AlexNet weights that cannot be fetched offline: not computed (None).
        Returns {"sr": {metric: mean}, "bicubic": {...}, "n": images} or None when the loader is empty."""
        self.generator.eval()
        acc = {"sr": [], "bicubic": []}
        for batch in loader:
            lr = batch[0].to(self.device, non_blocking=True)
            gt = batch[1].to(self.device, non_blocking=True)
            rec = self.generator(lr).float()
            acc["sr"].append(eval_metrics_u8(rec, gt, self.scale_factor))
            if len(batch) > 2 and torch.is_tensor(batch[2]):
                acc["bicubic"].append(eval_metrics_u8(batch[2].to(self.device, non_blocking=True), gt, self.scale_factor))
            if save_dir is not None:                                                 # :1506-1509
                for i in range(rec.shape[0]):
                    name = os.path.splitext(os.path.basename(str(batch[-1][i])))[0]
                    d = os.path.join(save_dir, 'validate', classname or '')
                    save_img1(rec[i], d, os.path.join(d, '%s_x%d_%d.png' % (name, self.scale_factor, epoch)))
        if not acc["sr"]:
            return None
        out = {"n": int(sum(m["mse"].numel() for m in acc["sr"]))}
        for who, ms in acc.items():
            if ms:
                cat = torch.stack([torch.cat([m[k] for m in ms]) for k in ("mse", "psnr", "ssim", "ergas")])   # one D2H copy
                sums = cat.sum(dim=1).tolist()
                out[who] = dict(zip(("mse", "psnr", "ssim", "ergas"), [v / cat.shape[1] for v in sums]))
        return out

    def _log_val(self, label, epoch, res, t0):
        """the 'val' log line of the reference (:1166-1185 / :1548-1565)"""
        if self.logger is None and _rank() == 0:
            self.logger = CsvLogger(os.path.join(self.save_dir, 'logs'))
        if self.logger is None:
            return
        rlt = OrderedDict(model=label, epoch=epoch, iters=epoch, time=time.time() - t0)
        for who, tag in (("bicubic", "bicubic"), ("sr", self.model_name.lower())):
            for k in ("mse", "psnr", "ssim", "ergas"):
                if who in res:
                    rlt['%s_%s' % (tag, k)] = res[who][k]
        self.logger.print_format_results('val', rlt)

    @torch.no_grad()
    def validate(self, epoch=0, mode='test', save_img=False):
        """reference :1058-1194.  Returns (psnr, ssim, ergas, lpips) of the SR output averaged over the test set (lpips is None:
        not computed offline), or None when there is nothing to validate on (synthetic training, missing test folder) — the
        caller's early-stopping heuristic then leaves its counters alone."""
        if self.generator is None:
            self.build(init=False)
            self.load_model()
        loader = self._eval_loader() if not self.synthetic_steps else None
        if loader is None:
            return None
        t0 = time.time()
        res = self._evaluate(loader)
        if res is None:
            return None
        self._log_val(self.model_name, epoch, res, t0)
        return res["sr"]["psnr"], res["sr"]["ssim"], res["sr"]["ergas"], None

    def mfeNew_validate(self, epoch=100, modelpath=None):
        self.build(init=False)
        if modelpath is not None:
            self.generator.load_state_dict(torch.load(modelpath, map_location=self.device), strict=False)   # :1270
            ops.bump_weight_generation()
        return self.validate(epoch=epoch, mode='test')

    @torch.no_grad()
    def mfeNew_validateByClass(self, epoch=100, save_img=False, modelpath=None):
        """reference :1393-1601: every class sub-folder of the first test dataset (UCMerced_LandUse's 21 land-use classes,
        :1430-1438) is evaluated separately — one 'val' log line per class, then the total over all images.
        Returns {class name: metrics, ..., "Total": metrics} with metrics = {"sr": {...}, "bicubic": {...}, "n": images}."""
        self.generator = self.new_generator()
        if modelpath is not None:
            self.generator.load_state_dict(torch.load(modelpath, map_location="cpu"), strict=False)          # :1401-1402
        self.generator.to(self.device).eval()
        ops.bump_weight_generation()
        root = os.path.join(self.data_dir, self.test_dataset[0])
        classes = [d for d in sorted(os.listdir(root)) if os.path.isdir(os.path.join(root, d))] if os.path.isdir(root) else []
        t0 = time.time()
        results, tot = OrderedDict(), None
        for cname in classes:
            loader = self._eval_loader([os.path.join(self.test_dataset[0], cname)])
            res = self._evaluate(loader, self.save_dir if save_img else None, epoch, cname) if loader is not None else None
            if res is None:
                continue
            results[cname] = res
            self._log_val(cname, epoch, res, t0)
            if tot is None:
                tot = {"n": 0, "sr": dict.fromkeys(res["sr"], 0.0), "bicubic": dict.fromkeys(res.get("bicubic", {}), 0.0)}
            tot["n"] += res["n"]
            for who in ("sr", "bicubic"):
                for k, v in res.get(who, {}).items():
                    tot[who][k] += v * res["n"]
        if tot is not None:
            for who in ("sr", "bicubic"):
                tot[who] = {k: v / tot["n"] for k, v in tot[who].items()}
            if not tot["bicubic"]:
                del tot["bicubic"]
            results["Total"] = tot
            self._log_val("Total", epoch, tot, t0)
        return results

    def mfe_test_single(self, img_fn, modelpath=None, tile=None, overlap=16):
        """reference :1603-1640: CenterCrop(test_crop_size) -> batch of `batch_size` identical copies ->
        G -> save [0] as uint8 (truncation), plus the PIL-bicubic baseline image (:1630,:1638).  `tile` (new): run large
        inputs as overlapped LR tiles of that size, which the reference cannot do because SGAM materialises an (HW)x(HW)
        attention.  Returns the SR image (3, H*s, W*s) on the device."""
        from PIL import Image
        import torchvision.transforms as transforms
        from ..data import pil_bicubic
        self.generator = self.new_generator()
        if modelpath is not None:
            self.generator.load_state_dict(torch.load(modelpath, map_location="cpu"), strict=False)   # :1612-1613
        self.generator.to(self.device).eval()
        ops.bump_weight_generation()
        img = transforms.Compose([transforms.CenterCrop(self.test_crop_size), transforms.ToTensor()])(Image.open(img_fn))
        input_img = img.unsqueeze(0).expand(self.batch_size, -1, -1, -1).contiguous().to(self.device)   # :1628-1629
        with torch.no_grad():
            if tile:
                recon = tiled_forward(self.generator, input_img[:1], self.scale_factor, tile, overlap)
            else:
                recon = self.generator(input_img)
        img_name = img_fn.split("/")[-1]
        save_img1(recon[0].float(), self.save_dir, os.path.join(self.save_dir, 'SR_%s_%s' % (self.model_name, img_name)))   # :1637
        # bicubic baseline (`img_interp`, utils/utils.py:755-783: ToPILImage -> PIL resize -> ToTensor), on the device
        u8 = (input_img[:1] * 255.0).clamp(0, 255).floor()
        h, w = u8.shape[-2:]
        bc = pil_bicubic(u8, h * self.scale_factor, w * self.scale_factor) / 255.0
        save_img1(bc[0], self.save_dir, os.path.join(self.save_dir, 'SR_Bicubic_%s' % img_name))                         # :1638
        return recon[0]

    # ------------------------------------------------------------------------------------------
    # checkpoints (reference :1197-1256) — plain state_dicts, NCHW fp32, reference key names
    # ------------------------------------------------------------------------------------------
    def save_epoch_network(self, save_dir, network, network_label, iter_label):
        os.makedirs(save_dir, exist_ok=True)
        path = os.path.join(save_dir, '{}_param_epoch_{}.pkl'.format(network_label, iter_label))
        torch.save({k: v.detach().cpu().clone() for k, v in network.state_dict().items()}, path)

    def load_epoch_network(self, load_path, network, strict=True):
        network.load_state_dict(torch.load(load_path, map_location=self.device), strict=strict)
        ops.bump_weight_generation()
        print('Trained model is loaded.')

    # ------------------------------------------------------------------------------------------
    # chain training (paper title; reference :716-721): the generator of scale s is warm-started from the trained
    # generator of the previous scale of the chain, the critic (whose input is always the 216^2 HR image) is carried over
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def warm_start(network, state_dict):
        """`load_state_dict(strict=False)` as the reference does (:720-721), with the one extension a chain across the two
        up-sampler families needs: entries whose SHAPE differs (x2/x4/x8 heads are 64->256, x3/x9 heads 64->576, SURVEY.md
        App. A) are skipped instead of raising, so they keep their fresh initialisation.  Returns (loaded, skipped) keys."""
        own = network.state_dict()
        ok = {k: v for k, v in state_dict.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
        skipped = [k for k in state_dict if k not in ok]
        network.load_state_dict(ok, strict=False)
        ops.bump_weight_generation()
        return sorted(ok), skipped

    def load_pretrained(self, G_path=None, D_path=None):
        for path, net in ((G_path, self.generator), (D_path, self.discriminator)):
            if path:
                sd = torch.load(path, map_location=self.device) if isinstance(path, str) else path
                _, skipped = self.warm_start(net, sd)
                print('Pretrained model is loaded (%d entries re-initialised: %s)' % (len(skipped), ", ".join(skipped[:4])))

    def chain_train(self, scales=(2, 3, 4), on_stage_end=None):
        """BASELINE.json configs[2]: train the scales of the chain one after the other in ONE process; every stage is a
        complete `train()` of the reference (fresh Adam state, `num_epochs` epochs, own save_dir/x<scale>) whose generator
        and critic start from the previous stage's final weights.  Returns {scale: (avg_loss_G, avg_loss_D)}."""
        base_dir, results, prev = self.save_dir, {}, None
        base_lr = self.lr                 # train()'s rollback heuristic halves self.lr: every stage starts from the configured rate
        for s in scales:
            self.lr = base_lr
            self.scale_factor = int(s)
            self.save_dir = os.path.join(base_dir, "x%d" % s)
            self.epoch = 0
            self._graph = None
            self._pack_plans = None
            self.pretrained_G, self.pretrained_D = (prev if prev is not None else (self.pretrained_G, self.pretrained_D))
            results[s] = self.train()
            prev = ({k: v.detach().clone() for k, v in self.generator.state_dict().items()},
                    {k: v.detach().clone() for k, v in self.discriminator.state_dict().items()})
            if on_stage_end is not None:
                on_stage_end(s, self)
        self.save_dir = base_dir
        return results

    def save_model(self, epoch=None):
        model_dir = os.path.join(self.save_dir, 'model')
        os.makedirs(model_dir, exist_ok=True)
        suffix = '_param_epoch_%d.pkl' % epoch if epoch is not None else '_param.pkl'
        torch.save({k: v.detach().cpu().clone() for k, v in self.generator.state_dict().items()}, model_dir + '/generator' + suffix)
        torch.save({k: v.detach().cpu().clone() for k, v in self.discriminator.state_dict().items()}, model_dir + '/discriminator' + suffix)
        print('Trained model is saved.')

    def load_model(self):
        name = os.path.join(self.save_dir, 'model') + '/generator_param.pkl'
        if os.path.exists(name):
            self.generator.load_state_dict(torch.load(name, map_location=self.device), strict=False)
            ops.bump_weight_generation()
            print('Trained model is loaded.')
            return True
        print('No model exists to load.')
        return False

    def load_epoch_model(self, epoch):
        name = os.path.join(self.save_dir, 'model') + '/generator_param_epoch_%d.pkl' % epoch
        if os.path.exists(name):
            self.generator.load_state_dict(torch.load(name, map_location=self.device))
            ops.bump_weight_generation()
            print('Trained model is loaded.')
            return True
        print('No model exists to load.')
        return False


def _rank():
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def tile_starts(size, tile, overlap):
    """start offsets of overlapped tiles covering [0, size)"""
    if size <= tile:
        return [0]
    step = tile - overlap
    starts = list(range(0, size - tile, step)) + [size - tile]
    return sorted(set(starts))


@torch.no_grad()
def tiled_forward(generator, lr, scale, tile, overlap=16, tile_batch=32):
    """Overlapped-tile inference (new): LR tiles of `tile`^2 with `overlap` LR pixels of
    overlap, feathered (linear ramp) blending of the SR tiles.  Parity is defined per tile: each tile's
    output equals the generator run on that tile alone.  Tiles (all the same shape) go through the generator
    `tile_batch` at a time."""
    n, c, h, w = lr.shape
    out = torch.zeros(n, c, h * scale, w * scale, device=lr.device, dtype=torch.float32)
    wsum = torch.zeros(1, 1, h * scale, w * scale, device=lr.device, dtype=torch.float32)
    th, tw = min(tile, h), min(tile, w)
    origins = [(y0, x0) for y0 in tile_starts(h, tile, overlap) for x0 in tile_starts(w, tile, overlap)]
    for i in range(0, len(origins), max(1, tile_batch // max(n, 1))):
        group = origins[i:i + max(1, tile_batch // max(n, 1))]
        batch = torch.cat([lr[:, :, y0:y0 + th, x0:x0 + tw] for y0, x0 in group], dim=0).contiguous()
        sr = generator(batch).float()
        for k, (y0, x0) in enumerate(group):
            wy = _feather(th * scale, overlap * scale, y0 > 0, y0 + th < h, lr.device)
            wx = _feather(tw * scale, overlap * scale, x0 > 0, x0 + tw < w, lr.device)
            wgt = wy.view(1, 1, -1, 1) * wx.view(1, 1, 1, -1)
            ys, xs = y0 * scale, x0 * scale
            out[:, :, ys:ys + th * scale, xs:xs + tw * scale] += sr[k * n:(k + 1) * n] * wgt
            wsum[:, :, ys:ys + th * scale, xs:xs + tw * scale] += wgt
    return out / wsum


def _feather(n, ramp, lo, hi, device):
    w = torch.ones(n, device=device)
    if ramp > 0:
        r = (torch.arange(ramp, device=device, dtype=torch.float32) + 1) / (ramp + 1)
        if lo:
            w[:ramp] = r
        if hi:
            w[n - ramp:] = r.flip(0)
    return w
