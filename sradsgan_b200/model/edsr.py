"""B200-native EDSR baseline — same class names, constructor / forward signatures and state_dict keys as the
reference's SRADSGAN/model/edsr.py (`Net` :23-75, trainer `EDSR` :78-955) and the two blocks it takes from
model/base_networks.py (`ConvBlock` :170-208, `ResnetBlock` :246-298), computed by the same sm_100a convolution
kernels as SRADSGAN (the second residual-conv workload, BASELINE.json configs[4]).

    x -> input_conv 3->256 -> 32 x [conv3x3 -> ReLU -> conv3x3 -> + x] -> mid_conv -> + skip
      -> [conv 256->1024 -> PixelShuffle(2) -> LeakyReLU(0.01)] x log2(scale) (ONE shared conv, like GAB_UP)
      -> output_conv 256->3

Each ResnetBlock is one autograd node (`ops.conv_act_conv`): ReLU in conv1's epilogue, the skip connection in
conv2's epilogue, ReLU' in the epilogue of conv2's input-gradient kernel.  The residual trunk stays fp32.
"""
import math
import os
import time
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import _lib, dp, ops
from .._lib import ACT_LRELU, ACT_NONE, ACT_RELU
from ..nn import Conv2d, LeakyReLU, PixelShuffle
from ..optim import FlatAdam
from ..utils import CsvLogger, psnr, save_img1, weights_init_normal
from .trainer import SRADSGAN, _rank, tiled_forward

_ACTS = {None: (ACT_NONE, 0.0), 'relu': (ACT_RELU, 0.0), 'lrelu': (ACT_LRELU, 0.2)}


class ConvBlock(nn.Module):
    """reference model/base_networks.py:170-208 — conv (+ activation); the normalisation variants are not used by EDSR."""

    def __init__(self, input_size, output_size, kernel_size=4, stride=2, padding=1, dilation=1, bias=True, activation=None, norm=None):
        super().__init__()
        if norm is not None:
            raise NotImplementedError("ConvBlock(norm=%r): EDSR builds every block with norm=None (model/edsr.py:27-60)" % (norm,))
        if activation not in _ACTS:
            raise NotImplementedError("ConvBlock(activation=%r): only None / 'relu' / 'lrelu' are built" % (activation,))
        self.conv = Conv2d(input_size, output_size, kernel_size, stride, padding, bias=bias, dilation=dilation)
        self.norm, self.activation = norm, activation

    def forward(self, x, out_dtype=None):
        act, slope = _ACTS[self.activation]
        return self.conv.fused(x, act, slope, out_dtype=out_dtype)


class ResnetBlock(nn.Module):
    """reference model/base_networks.py:246-298 with norm=None: conv1 -> act -> conv2 -> + x"""

    def __init__(self, num_filter, kernel_size=3, stride=1, padding=1, bias=True, activation='relu', norm='batch'):
        super().__init__()
        if norm is not None:
            raise NotImplementedError("ResnetBlock(norm=%r): EDSR uses norm=None (model/edsr.py:31)" % (norm,))
        if activation not in _ACTS:
            raise NotImplementedError("ResnetBlock(activation=%r): only None / 'relu' / 'lrelu' are built" % (activation,))
        self.conv1 = Conv2d(num_filter, num_filter, kernel_size, stride, padding, bias=bias)
        self.conv2 = Conv2d(num_filter, num_filter, kernel_size, stride, padding, bias=bias)
        self.norm, self.activation = norm, activation

    def forward(self, x):
        act, slope = _ACTS[self.activation]
        c1, c2 = self.conv1, self.conv2
        if c1.kernel_size == 3 and c1.stride == 1 and c1.padding == 1 and c1.bias is not None and c2.bias is not None:
            return ops.conv_act_conv(x, c1, c2, act, slope, residual=x, out_dtype=torch.float32)
        return c2.fused(c1.fused(x, act, slope), residual=x, out_dtype=torch.float32)


class Net(nn.Module):
    """EDSR generator (reference model/edsr.py:23-75).  As in the reference the up-sampling convolutions are hard-wired to
    256 channels (:42-48), so base_filter must be 256, and the stages of x4 / x8 / x9 share ONE convolution."""

    def __init__(self, num_channels, base_filter, num_residuals, upscale_factor=3):
        super().__init__()
        self.input_conv = ConvBlock(num_channels, base_filter, 3, 1, 1, activation=None, norm=None)
        self.residual_layers = nn.Sequential(*[ResnetBlock(base_filter, norm=None) for _ in range(num_residuals)])
        self.mid_conv = ConvBlock(base_filter, base_filter, 3, 1, 1, activation=None, norm=None)
        upsampling = []
        two = [Conv2d(256, 256 * 4, 3, 1, 1), PixelShuffle(2), LeakyReLU(inplace=True)]
        three = [Conv2d(256, 256 * 9, 3, 1, 1), PixelShuffle(3), LeakyReLU(inplace=True)]
        if (upscale_factor & (upscale_factor - 1)) == 0:
            for _ in range(int(math.log(upscale_factor, 2))):
                upsampling += two
        elif upscale_factor % 3 == 0:
            for _ in range(int(math.log(upscale_factor, 3))):
                upsampling += three
        self.upsampling = nn.Sequential(*upsampling)
        if len(upsampling) > 3:            # one conv applied at several stages: its gradient is a sum over the uses
            for p in upsampling[0].parameters():
                p._sr_shared = True
        self.output_conv = ConvBlock(base_filter, num_channels, 3, 1, 1, activation=None, norm=None)

    def weight_init(self, mean=0.0, std=0.02):
        for m in self.modules():
            weights_init_normal(m, mean=mean, std=std)

    def forward(self, x):
        x = ops.to_compute(x)
        out = self.input_conv(x, out_dtype=torch.float32)
        residual = out
        out = self.residual_layers(out)
        out = self.mid_conv.conv.fused(out, residual=residual, out_dtype=torch.float32)       # mid_conv + torch.add (:70-71)
        mods = list(self.upsampling)
        for i in range(0, len(mods), 3):   # conv -> PixelShuffle(r) -> LeakyReLU fused into one kernel
            out = mods[i].fused(out, ACT_LRELU, mods[i + 2].negative_slope, shuffle_r=mods[i + 1].upscale_factor)
        return self.output_conv(out, out_dtype=torch.float32)


class EDSR(SRADSGAN):
    """Trainer with the entry points of the reference's `EDSR` class (model/edsr.py:78-955): `train()`, `validate()`,
    `mfeNew_validate[ByClass]()`, `mfe_test_single()`, `save/load_*`.  One iteration (:246-265) = generator forward,
    L1 (or L2) pixel loss, backward, Adam — no discriminator, no VGG."""

    num_residuals = 32

    def new_generator(self):
        return Net(num_channels=self.num_channels, base_filter=256, num_residuals=self.num_residuals,
                   upscale_factor=self.scale_factor)                                   # :157

    def build(self, init=True):
        torch.manual_seed(self.seed)
        self.generator = self.new_generator()
        if init and self.epoch == 0:
            self.generator.apply(weights_init_normal)                                   # :181
        self.generator.to(self.device)
        self.optimizer_G = FlatAdam(self.generator, lr=self.lr, betas=(self.b1, self.b2))   # :184
        dp.broadcast_parameters(self.optimizer_G)
        self.reducer_G = dp.BucketReducer(self.optimizer_G, overlap=False)

    def _g_phase(self, imgs_lr, imgs_hr):
        self._repack()
        self.optimizer_G.zero_grad()                                                    # :252
        gen_hr = self.generator(imgs_lr)                                                # :255
        loss_G = self.criterion_content(gen_hr, imgs_hr)                                # :257-260
        self.reducer_G.arm()
        loss_G.backward()                                                               # :264
        ops.wgrad_join()
        return {"loss_G": loss_G.detach(), "gen_hr": gen_hr.detach()}

    def train_step(self, imgs_lr, imgs_hr, fuse_gp_backward=True):
        out = self._g_phase(imgs_lr, imgs_hr)
        scale = self.reducer_G.finish()
        self.optimizer_G.step(grad_scale=scale)                                         # :265
        out["loss_D"] = torch.zeros_like(out["loss_G"])                                 # the reference logs D_loss = 0 (:326)
        return out

    # -- CUDA-graph replay (same contract as SRADSGAN.graphed_step) --------------------------------
    def graphed_step(self, imgs_lr, imgs_hr):
        key = (tuple(imgs_lr.shape), tuple(imgs_hr.shape), self.optimizer_G.param_groups[0]["lr"])
        if self._graph is None or self._graph["key"] != key:
            self._capture(imgs_lr, imgs_hr, key)
        g = self._graph
        g["lr"].copy_(imgs_lr, non_blocking=True)
        g["hr"].copy_(imgs_hr, non_blocking=True)
        if len(g["graphs"]) == 1:
            g["graphs"][0].replay()
        else:
            g["graphs"][0].replay()
            dp.all_reduce_flat(self.optimizer_G.flat_grad)
            g["graphs"][1].replay()
        self.optimizer_G.step_count += 1
        self.optimizer_G.touch()          # masters changed through raw pointers: eager users of cached packed operands re-pack
        return g["out"]

    def _capture(self, imgs_lr, imgs_hr, key):
        world = dp.world_size()
        st = {"key": key, "lr": torch.empty(imgs_lr.shape, dtype=torch.float32, device=self.device),
              "hr": torch.empty(imgs_hr.shape, dtype=torch.float32, device=self.device)}
        st["lr"].copy_(imgs_lr); st["hr"].copy_(imgs_hr)
        oG = self.optimizer_G
        state = [oG.flat_param, oG.exp_avg, oG.exp_avg_sq, oG.step_t]
        snap = [t.clone() for t in state]
        red = self.reducer_G
        self.reducer_G = dp.NullReducer(world)
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    self.train_step(st["lr"], st["hr"])
            torch.cuda.current_stream().wait_stream(s)
            if os.environ.get("SR_PACK_PLAN", "1") == "1":
                self._pack_plans = [ops.PackPlan(oG.params)]
            n0 = _lib.backend().launch_count()
            if world == 1:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    st["out"] = self.train_step(st["lr"], st["hr"])
                graphs = [graph]
            else:
                g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, capture_error_mode="thread_local"):
                    out = self._g_phase(st["lr"], st["hr"])
                with torch.cuda.graph(g2, pool=g1.pool(), capture_error_mode="thread_local"):
                    oG.step(grad_scale=1.0 / world)
                out["loss_D"] = torch.zeros_like(out["loss_G"])
                st["out"] = out
                graphs = [g1, g2]
            st["launches"] = _lib.backend().launch_count() - n0
            torch.cuda.synchronize()
        finally:
            self.reducer_G = red
        for t, c in zip(state, snap):
            t.copy_(c)
        oG.step_count -= 3
        ops.bump_weight_generation()
        st["graphs"] = graphs
        st["pack_plans"] = getattr(self, "_pack_plans", None)
        self._graph = st

    # -- training loop (reference :150-390) ----------------------------------------------------------
    def train(self):
        self.build()
        model_dir = os.path.join(self.save_dir, 'model')
        os.makedirs(model_dir, exist_ok=True)
        if self.epoch != 0:                                                             # :176-179
            self.load_epoch_network(model_dir + '/generator_param_epoch_%d.pkl' % self.epoch, self.generator, strict=True)
        # the epoch loop (:207-372: staging copies, iteration, logging, validation, per-epoch checkpoint, the no-improvement
        # rollback + LR halving) is the SRADSGAN trainer's, run on this class's step (graph replay by default)
        avg_loss_G, _ = self._fit(model_dir, {"generator": self.generator})
        if _rank() == 0:
            self.save_model(epoch=None)
        return avg_loss_G

    def save_model(self, epoch=None):                                                    # :545-554 (generator only)
        model_dir = os.path.join(self.save_dir, 'model')
        os.makedirs(model_dir, exist_ok=True)
        suffix = '_param_epoch_%d.pkl' % epoch if epoch is not None else '_param.pkl'
        torch.save({k: v.detach().cpu().clone() for k, v in self.generator.state_dict().items()}, model_dir + '/generator' + suffix)
        print('Trained model is saved.')

    def mfe_test_single(self, img_fn, modelpath=None, tile=None, overlap=16):
        """reference model/edsr.py:921-955 — same contract as SRADSGAN.mfe_test_single; output name SR_EDSR_<file>."""
        from PIL import Image
        import torchvision.transforms as transforms
        self.generator = self.new_generator()
        if modelpath is not None:
            self.generator.load_state_dict(torch.load(modelpath, map_location="cpu"), strict=False)
        self.generator.to(self.device).eval()
        img = transforms.Compose([transforms.CenterCrop(self.test_crop_size), transforms.ToTensor()])(Image.open(img_fn))
        input_img = img.unsqueeze(0).expand(self.batch_size, -1, -1, -1).contiguous().to(self.device)
        with torch.no_grad():
            recon = tiled_forward(self.generator, input_img[:1], self.scale_factor, tile, overlap) if tile else self.generator(input_img)
        out_path = os.path.join(self.save_dir, 'SR_EDSR_%s' % img_fn.split("/")[-1])
        save_img1(recon[0].float().cpu(), self.save_dir, out_path)
        return recon[0]
