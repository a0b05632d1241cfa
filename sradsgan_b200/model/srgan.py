"""B200-native SRGAN sibling — same class names, constructor / forward signatures and state_dict keys as the reference's
SRADSGAN/model/srgan.py (`ResidualBlock` :57-71, `GeneratorResNet` :73-123, `Discriminator` :125-156, trainer `SRGAN` :158-1038),
on the kernels of the SRADSGAN hot path (the first of the sibling GANs).

    x -> conv 9x9 3->64 + ReLU -> 16 x [conv3x3 -> BN -> ReLU -> conv3x3 -> BN -> + x] -> conv3x3 -> BN -> + skip
      -> [conv 64->256 (576) -> BN -> PixelShuffle(2 | 3) -> ReLU] x stages (ONE shared conv / BN pair) -> conv 9x9 64->3 -> tanh

What runs where: every 3x3 convolution on the tcgen05 halo kernel (forward, input and weight gradients), train-mode
BatchNorm(+ReLU) on the fused `bn_leaky_relu` kernels (slope 0 = ReLU, slope 1 = identity), the critic on the fused conv /
BatchNorm Functions of the SRADSGAN critic (no attention, no gradient penalty), VGG19[:12] + max-pool, the MSE losses
(`ops.diff_mean_loss(p=2)`) and the fused Adam on the library kernels; the two 9x9 convolutions take the SIMT implicit-GEMM
kernel, and the residual adds / PixelShuffle copy / tanh stay on ATen (a sibling baseline, not the hot path).
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import dp, ops
from .._lib import ACT_NONE, ACT_RELU
from ..nn import BatchNorm2d, Conv2d, LeakyReLU, PixelShuffle, ReLU
from ..optim import FlatAdam
from ..utils import weights_init_normal
from .sradsgan import Discriminator as _AttnDiscriminator
from .sradsgan import FeatureExtractor  # noqa: F401  (reference model/srgan.py:44-55: the same VGG19[:12] extractor)
from .trainer import SRADSGAN, _rank


def _bn(x, bn, slope):
    """train mode: the fused BatchNorm + (Leaky)ReLU kernels (slope 1 = plain BatchNorm); eval mode: the module"""
    if bn.training:
        return ops.bn_leaky_relu(x, bn, slope)
    y = bn(x)
    return y if slope == 1.0 else F.leaky_relu(y, slope)


class ResidualBlock(nn.Module):
    """reference model/srgan.py:57-71"""

    def __init__(self, in_features):
        super().__init__()
        self.conv_block = nn.Sequential(Conv2d(in_features, in_features, 3, 1, 1), BatchNorm2d(in_features), ReLU(),
                                        Conv2d(in_features, in_features, 3, 1, 1), BatchNorm2d(in_features))

    def forward(self, x):
        cb = self.conv_block
        h = _bn(cb[0].fused(x), cb[1], 0.0)                       # conv -> BN -> ReLU
        return x + _bn(cb[3].fused(h), cb[4], 1.0).to(x.dtype)    # conv -> BN, + x   (:70; the trunk stays fp32)


class GeneratorResNet(nn.Module):
    """reference model/srgan.py:73-123"""

    def __init__(self, in_channels=3, out_channels=3, n_residual_blocks=16, upscale_factor=3):
        super().__init__()
        self.conv1 = nn.Sequential(Conv2d(in_channels, 64, 9, 1, 4), ReLU(inplace=True))
        self.res_blocks = nn.Sequential(*[ResidualBlock(64) for _ in range(n_residual_blocks)])
        self.conv2 = nn.Sequential(Conv2d(64, 64, 3, 1, 1), BatchNorm2d(64))
        upsampling = []
        two = [Conv2d(64, 64 * 4, 3, 1, 1), BatchNorm2d(64 * 4), PixelShuffle(2), ReLU(inplace=True)]
        three = [Conv2d(64, 64 * 9, 3, 1, 1), BatchNorm2d(64 * 9), PixelShuffle(3), ReLU(inplace=True)]
        if (upscale_factor & (upscale_factor - 1)) == 0:
            for _ in range(int(math.log(upscale_factor, 2))):
                upsampling += two                                  # the SAME module objects per stage, like the reference (:94-107)
        elif upscale_factor % 3 == 0:
            for _ in range(int(math.log(upscale_factor, 3))):
                upsampling += three
        self.upsampling = nn.Sequential(*upsampling)
        if len(upsampling) > 4:            # one conv / BatchNorm applied at several stages: their gradients are sums over the uses
            for m in upsampling[:2]:
                for p in m.parameters():
                    p._sr_shared = True
        self.conv3 = nn.Sequential(Conv2d(64, out_channels, 9, 1, 4), nn.Tanh())

    def forward(self, x):
        x = ops.to_compute(x)
        out1 = self.conv1[0].fused(x, ACT_RELU, 0.0, out_dtype=torch.float32)          # :116
        out = self.res_blocks(out1)
        out2 = _bn(self.conv2[0].fused(out), self.conv2[1], 1.0)                        # :118
        out = out1 + out2.to(out1.dtype)                                                # :119
        mods = list(self.upsampling)
        for i in range(0, len(mods), 4):   # conv -> BN -> PixelShuffle -> ReLU; ReLU commutes with the shuffle: fused into the BN kernel
            out = F.pixel_shuffle(_bn(mods[i].fused(out), mods[i + 1], 0.0), mods[i + 2].upscale_factor)
        return torch.tanh(self.conv3[0].fused(out, ACT_NONE, 0.0, out_dtype=torch.float32))   # :121


class Discriminator(_AttnDiscriminator):
    """reference model/srgan.py:125-156: the strided-conv BatchNorm critic without the attention pair (same Sequential indices)"""

    def __init__(self, in_channels=3):
        super().__init__(in_channels=in_channels, attention=False)


class SRGAN(SRADSGAN):
    """Trainer with the entry points of the reference's `SRGAN` class (model/srgan.py:158-1038).  One iteration (:343-381):
    G: MSE(gen, hr) + 6e-3 MSE(VGG(gen), VGG(hr)) + 1e-3 MSE(D(gen), 1);  D: (MSE(D(hr), 1) + MSE(D(gen.detach()), 0)) / 2;
    Adam on both, no gradient penalty, no weight clamp.  Data parallel / CUDA-graph replay / validation come from the base class."""

    n_residual_blocks = 16

    def new_generator(self):
        return GeneratorResNet(in_channels=3, out_channels=3, n_residual_blocks=self.n_residual_blocks,
                               upscale_factor=self.scale_factor)                                          # :256

    def build(self, init=True):
        torch.manual_seed(self.seed)
        self.generator = self.new_generator()
        self.discriminator = Discriminator()                                                               # :257
        vsd = torch.load(self.vgg_state, map_location="cpu") if isinstance(self.vgg_state, str) else self.vgg_state
        self.feature_extractor = FeatureExtractor(state_dict=vsd)                                          # :258
        if init and self.epoch == 0:
            self.generator.apply(weights_init_normal)                                                      # :279-280
            self.discriminator.apply(weights_init_normal)
        for m in (self.generator, self.discriminator, self.feature_extractor):
            m.to(self.device)
        self.optimizer_G = FlatAdam(self.generator, lr=self.lr, betas=(self.b1, self.b2))                  # :274
        self.optimizer_D = FlatAdam(self.discriminator, lr=self.lr, betas=(self.b1, self.b2))              # :275 (no clamp)
        dp.broadcast_parameters(self.optimizer_G)
        dp.broadcast_parameters(self.optimizer_D)
        self.reducer_G = dp.BucketReducer(self.optimizer_G, overlap=False)
        self.reducer_D = dp.BucketReducer(self.optimizer_D, overlap=False)

    @staticmethod
    def _mse_to(pred, value):
        """nn.MSELoss()(pred, full_like(pred, value)) — criterion_GAN against the `valid` / `fake` patch tensors (:287-288)"""
        return ops.diff_mean_loss(pred, torch.full_like(pred, value), 2)

    def _g_phase(self, imgs_lr, imgs_hr):
        G, D, Fx = self.generator, self.discriminator, self.feature_extractor
        self._repack()
        imgs_lr, imgs_hr = ops.to_compute(imgs_lr), ops.to_compute(imgs_hr)
        self.optimizer_G.zero_grad()                                                     # :343
        for p in self.optimizer_D.params:
            p.requires_grad_(False)          # D's weight gradients of the G step are discarded by the reference (:367 zero_grad)
        gen_hr = G(imgs_lr)                                                              # :346
        loss_gan = self._mse_to(D(gen_hr), 1.0)                                          # :348-349
        with torch.no_grad():
            real_features = Fx(imgs_hr)                                                  # :353
        loss_content = ops.diff_mean_loss(Fx(gen_hr), real_features, 2)                  # :352-354
        mse = ops.diff_mean_loss(gen_hr, imgs_hr, 2)                                     # :358
        loss_G = mse + 6e-3 * loss_content + 1e-3 * loss_gan                             # :361
        self.reducer_G.arm()
        loss_G.backward()
        ops.wgrad_join()
        for p in self.optimizer_D.params:
            p.requires_grad_(True)
        return {"loss_G": loss_G.detach(), "pixel": mse.detach(), "content": loss_content.detach(), "adv": loss_gan.detach(),
                "gen_hr": gen_hr.detach(), "_hr_nhwc": imgs_hr}

    def _d_phase(self, imgs_hr, gen_det, fuse_gp_backward=True):
        D = self.discriminator
        self.optimizer_D.zero_grad()                                                     # :367
        loss_real = self._mse_to(D(imgs_hr), 1.0)                                        # :370
        loss_fake = self._mse_to(D(gen_det), 0.0)                                        # :371
        loss_D = (loss_real + loss_fake) / 2                                             # :377
        loss_D.backward()
        ops.wgrad_join()
        return {"loss_D": loss_D.detach(), "gp": torch.zeros((), device=imgs_hr.device)}

    def mfe_test_single(self, img_fn, modelpath=None, tile=None, overlap=16):
        raise NotImplementedError("SRGAN: single-image test entry point is not built (train / validate only)")
