"""B200-native NDSRGAN sibling — same class names, constructor / forward signatures and state_dict keys as the reference's
SRADSGAN/model/ndsrgan.py (`CL` :57-58, `DenseBlock` :60-77, `DCRDB` :79-93, `DRRDBnet` :95-169, `GeneratorResNet` :171-223,
`Discriminator` :225-258, trainer `NDSRGAN` :260-) on the library's kernels.

    x -> conv 3->64 -> 23 densely connected DCRDBs (each: three dense blocks of four 3x3 conv+LeakyReLU(0.2) layers with 32-channel
         growth and a 192->64 fusion conv, plus a 64->64 conv; every skip scaled by 0.2) -> conv -> + skip
      -> [nearest x2 | x3 -> conv 64->64 -> LeakyReLU(0.2)] x stages (ONE shared conv) -> conv + LeakyReLU(0.2) -> conv 64->3
    critic: four 4x4 conv blocks (stride 2, 2, 2, 1; BatchNorm on all but the first) + a 4x4 output conv

What runs where: the 64- / 128- / 192-channel -> 64 convolutions (dense-block fusion convs — stacked-tap mode —, DCRDB convs, trunk and
up-sampler convs) on the tcgen05 halo kernel; the growth layers (32 output channels, 64 / 96 / 128 / 160 input channels) and the 4x4
critic convolutions on the SIMT implicit-GEMM kernel of the same library (forward, input and weight gradients; no tensor-core path for
those shapes); BatchNorm + LeakyReLU, VGG19[:12], the fused Adam on the library kernels; channel concatenation, the scaled skip sums,
nearest up-sampling and the Smooth-L1 criteria stay on ATen (a sibling baseline, not the hot path).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import dp, ops
from .._lib import ACT_LRELU, ACT_NONE
from ..nn import BatchNorm2d, Conv2d, LeakyReLU
from ..optim import FlatAdam
from ..utils import weights_init_normal
from .sradsgan import Discriminator as _BaseDiscriminator
from .sradsgan import FeatureExtractor  # noqa: F401
from .trainer import SRADSGAN


def CL(in_channels, out_channels):
    """reference model/ndsrgan.py:57-58"""
    return nn.Sequential(Conv2d(in_channels, out_channels, 3, 1, 1), LeakyReLU(0.2, inplace=True))


class DenseBlock(nn.Module):
    """reference model/ndsrgan.py:60-77"""

    def __init__(self, nf, nc, CL_num=4):
        super().__init__()
        self.CL_blocks = nn.Sequential(*[CL(nc * j + nf, nc) for j in range(CL_num)])
        self.conv = Conv2d(nc * CL_num + nf, nf, 3, 1, 1)

    def forward(self, x):
        out1 = x.float()
        x = ops.to_compute(x)
        for blk in self.CL_blocks:
            y = blk[0].fused(x, ACT_LRELU, blk[1].negative_slope)
            x = torch.cat((x, y), dim=1)
        return out1 + self.conv.fused(x, out_dtype=torch.float32) * 0.2


class DCRDB(nn.Module):
    """reference model/ndsrgan.py:79-93"""

    def __init__(self, nf, nc):
        super().__init__()
        self.RDB1 = DenseBlock(nf, nc)
        self.RDB2 = DenseBlock(nf, nc)
        self.RDB3 = DenseBlock(nf, nc)
        self.conv = Conv2d(nf, nf, 3, 1, 1)

    def forward(self, x):
        x = x.float()
        out1 = self.RDB1(x)
        out2 = self.RDB2(x + 0.2 * out1)
        out3 = self.RDB3(x + 0.2 * out1 + 0.2 * out2)
        out4 = self.conv.fused(x + 0.2 * out1 + 0.2 * out2 + 0.2 * out3, out_dtype=torch.float32)
        return out4 * 0.2 + x


class DRRDBnet(nn.Module):
    """reference model/ndsrgan.py:95-169: 23 DCRDBs, block k fed x + 0.2 * (sum of the outputs of blocks 1 .. k-1).
    `n_blocks` (new, default 23 = the reference's fixed count) only exists so that tests can build a shorter trunk."""

    def __init__(self, nf, nc, n_blocks=23):
        super().__init__()
        self.n_blocks = n_blocks
        for k in range(1, n_blocks + 1):
            setattr(self, "DRRDB%d" % k, DCRDB(nf, nc))

    def forward(self, x):
        acc = x.float()                      # x + 0.2 * m1 + ... accumulated in the reference's left-to-right order
        for k in range(1, self.n_blocks + 1):
            acc = acc + 0.2 * getattr(self, "DRRDB%d" % k)(acc)
        return acc


class GeneratorResNet(nn.Module):
    """reference model/ndsrgan.py:171-223"""

    def __init__(self, in_channels=3, out_channels=3, nf=64, nc=32, upscale_factor=3, n_blocks=23):
        super().__init__()
        self.conv1 = nn.Sequential(Conv2d(in_channels, nf, 3, 1, 1))
        self.DCRDB_block = DRRDBnet(nf=nf, nc=nc, n_blocks=n_blocks)
        self.conv2 = Conv2d(nf, nf, 3, 1, 1)
        upsampling = []
        two = [nn.UpsamplingNearest2d(scale_factor=2), Conv2d(nf, nf, 3, 1, 1), LeakyReLU(0.2, inplace=True)]
        three = [nn.UpsamplingNearest2d(scale_factor=3), Conv2d(nf, nf, 3, 1, 1), LeakyReLU(0.2, inplace=True)]
        if (upscale_factor & (upscale_factor - 1)) == 0:
            for _ in range(int(math.log(upscale_factor, 2))):
                upsampling += two                                  # the SAME module objects per stage, like the reference (:199-204)
        elif upscale_factor % 3 == 0:
            for _ in range(int(math.log(upscale_factor, 3))):
                upsampling += three
        self.upsampling = nn.Sequential(*upsampling)
        if len(upsampling) > 3:
            for p in upsampling[1].parameters():
                p._sr_shared = True
        self.conv3 = nn.Sequential(Conv2d(nf, nf, 3, 1, 1), LeakyReLU(0.2, inplace=True), Conv2d(nf, out_channels, 3, 1, 1))

    def forward(self, x):
        x = ops.to_compute(x)
        out = self.conv1[0].fused(x, out_dtype=torch.float32)                                      # :217
        trunk = self.conv2.fused(self.DCRDB_block(out), out_dtype=torch.float32)                   # :218
        out = out + trunk                                                                          # :219
        mods = list(self.upsampling)
        for i in range(0, len(mods), 3):   # nearest up-sampling (ATen) -> conv + LeakyReLU (one kernel)
            out = mods[i + 1].fused(F.interpolate(ops.to_compute(out), scale_factor=mods[i].scale_factor, mode="nearest"),
                                    ACT_LRELU, mods[i + 2].negative_slope)
        out = self.conv3[0].fused(out, ACT_LRELU, self.conv3[1].negative_slope)
        return self.conv3[2].fused(out, ACT_NONE, 0.0, out_dtype=torch.float32)                    # :221-222


class Discriminator(_BaseDiscriminator):
    """reference model/ndsrgan.py:225-258: four 4x4 conv blocks + a 4x4 output conv; forward / fused-block logic of the base critic"""

    def __init__(self, in_channels=3):
        nn.Module.__init__(self)
        layers, in_filters = [], in_channels
        for out_filters, stride, normalize in [(64, 2, False), (128, 2, True), (256, 2, True), (512, 1, True)]:
            layers.append(Conv2d(in_filters, out_filters, 4, stride, 1))
            if normalize:
                layers.append(BatchNorm2d(out_filters))
            layers.append(LeakyReLU(0.2, inplace=True))
            in_filters = out_filters
        layers.append(Conv2d(out_filters, 1, 4, 1, 1))
        self.model = nn.Sequential(*layers)
        self.block_taps = None


class NDSRGAN(SRADSGAN):
    """Trainer with the entry points of the reference's `NDSRGAN` class.  One iteration (model/ndsrgan.py:414-456), every criterion a
    Smooth-L1 loss (:325-329): G: 1e-2 SL1(gen, hr) + SL1(VGG(gen), VGG(hr)) + 2.5e-3 SL1(D(gen), 1);
    D: (SL1(D(hr), 1) + SL1(D(gen.detach()), 0)) / 2; Adam on both, no gradient penalty, no weight clamp."""

    n_blocks = 23

    def new_generator(self):
        return GeneratorResNet(in_channels=3, out_channels=3, nf=64, nc=32, upscale_factor=self.scale_factor, n_blocks=self.n_blocks)   # :319

    def build(self, init=True):
        torch.manual_seed(self.seed)
        self.generator = self.new_generator()
        self.discriminator = Discriminator()                                                               # :320
        vsd = torch.load(self.vgg_state, map_location="cpu") if isinstance(self.vgg_state, str) else self.vgg_state
        self.feature_extractor = FeatureExtractor(state_dict=vsd)                                          # :321
        if init and self.epoch == 0:
            self.generator.apply(weights_init_normal)                                                      # :344-345
            self.discriminator.apply(weights_init_normal)
        for m in (self.generator, self.discriminator, self.feature_extractor):
            m.to(self.device)
        self.optimizer_G = FlatAdam(self.generator, lr=self.lr, betas=(self.b1, self.b2))                  # :348
        self.optimizer_D = FlatAdam(self.discriminator, lr=self.lr, betas=(self.b1, self.b2))              # :349 (no clamp)
        dp.broadcast_parameters(self.optimizer_G)
        dp.broadcast_parameters(self.optimizer_D)
        self.reducer_G = dp.BucketReducer(self.optimizer_G, overlap=False)
        self.reducer_D = dp.BucketReducer(self.optimizer_D, overlap=False)

    @staticmethod
    def _sl1(a, b):
        """torch.nn.SmoothL1Loss() (beta = 1, mean) in fp32 — ATen: the library has L1 / L2 reductions only"""
        return F.smooth_l1_loss(a.float(), b.float())

    def _g_phase(self, imgs_lr, imgs_hr):
        G, D, Fx = self.generator, self.discriminator, self.feature_extractor
        self._repack()
        imgs_lr, imgs_hr = ops.to_compute(imgs_lr), ops.to_compute(imgs_hr)
        self.optimizer_G.zero_grad()                                                     # :414
        for p in self.optimizer_D.params:
            p.requires_grad_(False)          # D's weight gradients of the G step are discarded by the reference (:441 zero_grad)
        gen_hr = G(imgs_lr)                                                              # :417
        validity = D(gen_hr)                                                             # :419
        loss_gan = self._sl1(validity, torch.ones_like(validity))                        # :420
        with torch.no_grad():
            real_features = Fx(imgs_hr)                                                  # :424
        loss_content = self._sl1(Fx(gen_hr), real_features)                              # :423-425
        pix = self._sl1(gen_hr, imgs_hr)                                                 # :429 (`mse_loss_G` is a Smooth-L1 loss too)
        loss_G = 1e-2 * pix + loss_content + 2.5e-3 * loss_gan                           # :432
        self.reducer_G.arm()
        loss_G.backward()
        ops.wgrad_join()
        for p in self.optimizer_D.params:
            p.requires_grad_(True)
        return {"loss_G": loss_G.detach(), "pixel": pix.detach(), "content": loss_content.detach(), "adv": loss_gan.detach(),
                "gen_hr": gen_hr.detach(), "_hr_nhwc": imgs_hr}

    def _d_phase(self, imgs_hr, gen_det, fuse_gp_backward=True):
        D = self.discriminator
        self.optimizer_D.zero_grad()                                                     # :441
        d_real, d_fake = D(imgs_hr), D(gen_det)
        loss_D = (self._sl1(d_real, torch.ones_like(d_real)) + self._sl1(d_fake, torch.zeros_like(d_fake))) / 2   # :444-451
        loss_D.backward()
        ops.wgrad_join()
        return {"loss_D": loss_D.detach(), "gp": torch.zeros((), device=imgs_hr.device)}

    def mfe_test_single(self, img_fn, modelpath=None, tile=None, overlap=16):
        raise NotImplementedError("NDSRGAN: single-image test entry point is not built (train / validate only)")
