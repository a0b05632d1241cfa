"""B200-native SRAGAN sibling — same class names, constructor / forward signatures and state_dict keys as the reference's
SRADSGAN/model/sragan.py (`GeneratorResNet` :147-237, `Discriminator` :239-277, trainer `SRAGAN` :279-) and the two blocks it takes
from model/base_networks.py (`BasicBlock` :958-1070, `ResidualBlock_Block_WithAttention` :1505-1594, both with norm_type=None),
on the kernels of the SRADSGAN hot path (SRAGAN is SRADSGAN's predecessor — 64 -> 64 -> 64 attention blocks instead
of the 64 -> 256 -> 64 RABs, no multi-scale block, no dense sampling, BatchNorm in conv2 and in the up-sampler, tanh output).

    x -> conv 3->64 + LeakyReLU(0.01) -> 12 x [ (n-1) x BasicBlock(lrelu) -> BasicBlock(no act) -> CA -> SA -> conv1x1 -> + x ]
      -> conv3x3 -> BN -> + skip -> CAM -> PAM -> conv1x1 -> [conv 64->256 (576) -> BN -> PixelShuffle -> LeakyReLU(0.01)] x stages
      -> conv 64->3 -> tanh
    BasicBlock: conv3x3 + LeakyReLU(0.2) -> conv3x3 -> CA -> SA -> conv1x1 -> + x -> LeakyReLU(0.2)

What runs where: the two 3x3 convolutions of a BasicBlock are one `ops.conv_act_conv` node on the tcgen05 halo kernel whose epilogue
emits the pooling partials, CA -> SA -> conv1x1 -> + x is the fused local-attention chain (one launch forward, three backward), CAM / PAM
are the CGAM / SGAM kernels, BatchNorm(+LeakyReLU) the fused `bn_leaky_relu` kernels; the critic, the losses, the gradient penalty and the
training iteration are those of SRADSGAN (the reference's step bodies are identical, model/sragan.py:642-705 vs model/sradsgan.py:829-892).
The activation behind a BasicBlock's residual add, the PixelShuffle copy and tanh stay on ATen.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .._lib import ACT_LRELU, ACT_NONE, ACT_RELU
from ..nn import BatchNorm2d, Conv2d, LeakyReLU, PixelShuffle
from .edsr import ConvBlock
from .sradsgan import CGAM as CAM_Module
from .sradsgan import SGAM as PAM_Module
from .sradsgan import Discriminator, FeatureExtractor, _fused_la, _la_forward, _la_init  # noqa: F401
from .srgan import _bn
from .trainer import SRADSGAN

_ACT = {'lrelu': (ACT_LRELU, 0.2), 'relu': (ACT_RELU, 0.0)}


class BasicBlock(nn.Module):
    """reference model/base_networks.py:958-1070 (norm_type=None): ConvBlock(act) -> ConvBlock -> local attention -> + x -> act"""
    expansion = 1

    def __init__(self, inplanes, planes, kernel_size=3, stride=1, padding=1, bias=True, dilation=1, norm_type='batch', act_type=None,
                 la_mode='CA-SA', pool_mode='Avg|Max', addconv=True, downsample=None):
        super().__init__()
        if norm_type is not None:
            raise NotImplementedError("BasicBlock(norm_type=%r): SRAGAN builds its blocks with norm_type=None (model/sragan.py:165)" % (norm_type,))
        if act_type not in (None, 'lrelu', 'relu'):
            raise NotImplementedError("BasicBlock(act_type=%r): only None / 'relu' / 'lrelu' are built" % (act_type,))
        if inplanes != planes:
            raise NotImplementedError("BasicBlock: inplanes != planes is not used by SRAGAN")
        self.inplanes, self.planes = inplanes, planes
        self.conv1 = ConvBlock(inplanes, planes, kernel_size, stride, padding, dilation=dilation, bias=bias, activation=act_type, norm=None)
        self.conv2 = ConvBlock(planes, planes, kernel_size, stride, padding, dilation=dilation, bias=bias, activation=None, norm=None)
        _la_init(self, planes, la_mode, pool_mode, addconv)
        self.act_type = act_type

    def forward(self, x):
        xc = ops.to_compute(x)
        c1, c2 = self.conv1.conv, self.conv2.conv
        if (self.act_type in _ACT and c1.kernel_size == 3 and c1.stride == 1 and c1.padding == 1 and c1.bias is not None
                and c2.bias is not None):
            act, slope = _ACT[self.act_type]
            out = ops.conv_act_conv(xc, c1, c2, act, slope, want_pool=_fused_la(self))      # both convolutions: one autograd node
        else:
            out = self.conv2(self.conv1(xc))
        z = _la_forward(self, out, x)                       # CA -> SA -> conv1x1 -> + x (fp32 residual stream)
        if self.act_type in _ACT:
            z = F.leaky_relu(z, _ACT[self.act_type][1])     # `out += residual; out = self.act(out)` (:1064-1068)
        return z


class ResidualBlock_Block_WithAttention(nn.Module):
    """reference model/base_networks.py:1505-1594: (n_blocks - 1) blocks + last_conv block -> local attention -> + x"""

    def __init__(self, block, n_blocks=1, nc=64, gc=32, kernel_size=3, stride=1, bias=True, padding=1, norm_type='batch',
                 act_type='relu', mode='CNA', rla_mode='CA-SA', bla_mode='CA-SA', pool_mode='Avg|Max', addconv=True):
        super().__init__()
        mk = lambda act: block(nc, nc, kernel_size=kernel_size, bias=bias, stride=stride, padding=padding, norm_type=norm_type,
                               act_type=act, la_mode=bla_mode, pool_mode=pool_mode, addconv=addconv)
        self.blocks = nn.Sequential(*[mk(act_type) for _ in range(n_blocks - 1)])
        self.last_conv = mk(None if mode == 'CNA' else act_type)
        _la_init(self, nc, rla_mode, pool_mode, addconv)

    def forward(self, x):
        out = self.last_conv(self.blocks(x))
        return _la_forward(self, ops.to_compute(out), x)


class GeneratorResNet(nn.Module):
    """reference model/sragan.py:147-237"""

    def __init__(self, buildingblock, in_channels=3, out_channels=3, n_residual_blocks=12, n_basic_blocks=1,
                 rla_mode='CA-SA', bla_mode='CA-SA', ga_mode='CA-SA', pool_mode='Avg|Max', addconv=True, upscale_factor=3):
        super().__init__()
        self.ga_mode, self.addconv = ga_mode, addconv
        self.conv1 = nn.Sequential(Conv2d(in_channels, 64, 3, 1, 1), LeakyReLU(inplace=True))
        self.res_blocks = nn.Sequential(*[
            buildingblock(BasicBlock, n_blocks=n_basic_blocks, nc=64, gc=32, kernel_size=3, stride=1, padding=1, norm_type=None,
                          act_type='lrelu', mode='CNA', rla_mode=rla_mode, bla_mode=bla_mode, pool_mode=pool_mode, addconv=addconv)
            for _ in range(n_residual_blocks)])
        self.conv2 = nn.Sequential(Conv2d(64, 64, 3, 1, 1), BatchNorm2d(64))
        if ga_mode.find('CA') != -1:
            self.ca = CAM_Module(64)
        if ga_mode.find('SA') != -1:
            self.sa = PAM_Module(64)
        if ga_mode.find('-') != -1 and addconv:
            self.conv = Conv2d(64, 64, 1, bias=True)
        if ga_mode.find('|') != -1:
            self.conv = Conv2d(64 * 2, 64, 1, bias=True)
        upsampling = []
        two = [Conv2d(64, 64 * 4, 3, 1, 1), BatchNorm2d(64 * 4), PixelShuffle(2), LeakyReLU(inplace=True)]
        three = [Conv2d(64, 64 * 9, 3, 1, 1), BatchNorm2d(64 * 9), PixelShuffle(3), LeakyReLU(inplace=True)]
        if (upscale_factor & (upscale_factor - 1)) == 0:
            for _ in range(int(math.log(upscale_factor, 2))):
                upsampling += two                                  # the SAME module objects per stage, like the reference (:191-204)
        elif upscale_factor % 3 == 0:
            for _ in range(int(math.log(upscale_factor, 3))):
                upsampling += three
        self.upsampling = nn.Sequential(*upsampling)
        if len(upsampling) > 4:
            for m in upsampling[:2]:
                for p in m.parameters():
                    p._sr_shared = True
        self.conv3 = nn.Sequential(Conv2d(64, out_channels, 3, 1, 1), nn.Tanh())

    def forward(self, x):
        x = ops.to_compute(x)
        out1 = self.conv1[0].fused(x, ACT_LRELU, self.conv1[1].negative_slope, out_dtype=torch.float32)    # :216
        out = self.res_blocks(out1)
        out2 = _bn(self.conv2[0].fused(out), self.conv2[1], 1.0)                                            # :218
        out = out1 + out2.to(out1.dtype)                                                                    # :219
        m = self.ga_mode
        if m == 'CA':
            out = self.ca(out)
        elif m == 'SA':
            out = self.sa(out)
        elif m in ('CA-SA', 'SA-CA'):
            out = self.sa(self.ca(out)) if m == 'CA-SA' else self.ca(self.sa(out))
            if self.addconv:
                out = self.conv.fused(out)
        elif m == 'CA|SA':
            out = self.conv.fused(torch.cat([ops.to_compute(self.ca(out)), ops.to_compute(self.sa(out))], dim=1))
        mods = list(self.upsampling)
        for i in range(0, len(mods), 4):   # conv -> BN -> PixelShuffle -> LeakyReLU; the activation commutes with the shuffle: fused into the BN kernel
            out = F.pixel_shuffle(_bn(mods[i].fused(out), mods[i + 1], mods[i + 3].negative_slope), mods[i + 2].upscale_factor)
        return torch.tanh(self.conv3[0].fused(out, ACT_NONE, 0.0, out_dtype=torch.float32))                 # :236


class SRAGAN(SRADSGAN):
    """Trainer with the entry points of the reference's `SRAGAN` class.  Its iteration (model/sragan.py:642-705) is SRADSGAN's
    (:829-892) line for line — L1 pixel + VGG + WGAN losses, WGAN-GP, Adam, critic clamp — so only the generator differs."""

    n_residual_blocks, n_basic_blocks = 12, 5

    def new_generator(self):
        return GeneratorResNet(ResidualBlock_Block_WithAttention, n_residual_blocks=self.n_residual_blocks,
                               n_basic_blocks=self.n_basic_blocks, rla_mode='CA-SA', bla_mode='CA-SA', ga_mode='CA-SA',
                               pool_mode='Avg|Max', upscale_factor=self.scale_factor)                        # model/sragan.py:465-466
