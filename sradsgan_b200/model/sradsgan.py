"""B200-native SRADSGAN networks — same class names, constructor signatures, forward signatures and
state_dict keys as the reference's SRADSGAN/model/sradsgan.py (cited per class), computed by the
sradsgan_b200 CUDA library (hand-written sm_100a kernels behind a C ABI; no cuDNN, no CPU fallback).

Layout: every module takes/returns logically-NCHW tensors; internally activations are NHWC
(torch.channels_last) in `ops.config.compute_dtype` (bf16 by default), parameters stay OIHW fp32.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Parameter

from .. import ops
from .._lib import ACT_LRELU, ACT_NONE, ACT_RELU
from ..nn import BatchNorm2d, Conv2d, LeakyReLU, MaxPool2d, PixelShuffle, ReLU


class GANLoss(nn.Module):
    """reference model/sradsgan.py:35-67"""

    def __init__(self, gan_type, real_label_val=1.0, fake_label_val=0.0):
        super().__init__()
        self.gan_type = gan_type.lower()
        self.real_label_val = real_label_val
        self.fake_label_val = fake_label_val
        if self.gan_type == 'vanilla':
            self.loss = nn.BCEWithLogitsLoss()
        elif self.gan_type == 'lsgan':
            self.loss = nn.MSELoss()
        elif self.gan_type == 'wgan-gp':
            self.loss = lambda inp, target: ops.mean_loss(inp, -1.0 if target else 1.0)
        else:
            raise NotImplementedError('GAN type [{:s}] is not found'.format(self.gan_type))

    def get_target_label(self, input, target_is_real):
        if self.gan_type == 'wgan-gp':
            return target_is_real
        return torch.empty_like(input).fill_(self.real_label_val if target_is_real else self.fake_label_val)

    def forward(self, input, target_is_real):
        return self.loss(input, self.get_target_label(input, target_is_real))


class FeatureExtractor(nn.Module):
    """VGG19 features[:12] (reference model/sradsgan.py:88-99).  The reference downloads torchvision's
    pretrained weights; offline, pass `state_dict=` (keys `feature_extractor.{0,2,5,7,10}.{weight,bias}`)
    or keep the seeded default init.  Parameters are frozen: the reference never optimises them and
    discards their gradients."""

    def __init__(self, state_dict=None):
        super().__init__()
        cfg = [(3, 64), (64, 64), 'M', (64, 128), (128, 128), 'M', (128, 256)]
        layers = []
        for c in cfg:
            if c == 'M':
                layers.append(MaxPool2d(2, 2))
            else:
                layers += [Conv2d(c[0], c[1], 3, 1, 1), ReLU(True)]
        self.feature_extractor = nn.Sequential(*layers)
        if state_dict is not None:
            self.load_state_dict(state_dict, strict=True)
        for p in self.parameters():
            p.requires_grad_(False)
            p._sr_frozen = True        # ops.packed: packed operands of frozen weights may be reused inside CUDA graphs

    def forward(self, img):
        out = img
        for m in self.feature_extractor:
            if isinstance(m, Conv2d):
                out = m.fused(out, ACT_RELU)
            elif isinstance(m, MaxPool2d):
                out = m(out)
        return out


def _mlp_gate(fc1, fc2, x):
    """sigmoid(W2 relu(W1 avg) + W2 relu(W1 max)) — the shared bias-free 1x1 'MLP' of CLAM."""
    n, c = x.shape[0], x.shape[1]
    w1 = fc1.weight.view(fc1.out_channels, c)
    w2 = fc2.weight.view(c, fc2.in_channels)
    avg = x.mean(dim=(2, 3), dtype=torch.float32)
    mx = F.adaptive_max_pool2d(x, 1).flatten(1).float()
    return avg, mx, w1, w2


class CLAM(nn.Module):
    """Channel local attention (reference model/sradsgan.py:101-127)."""

    def __init__(self, in_planes, ratio=16, pool_mode='Avg|Max'):
        super().__init__()
        self.pool_mode = pool_mode
        self.fc1 = Conv2d(in_planes, in_planes // ratio, 1, bias=False)
        self.fc2 = Conv2d(in_planes // ratio, in_planes, 1, bias=False)

    def forward(self, x):
        x = ops.to_compute(x)
        avg, mx, w1, w2 = _mlp_gate(self.fc1, self.fc2, x)
        mlp = lambda v: F.relu(v @ w1.t()) @ w2.t()
        if self.pool_mode == 'Avg':
            o = mlp(avg)
        elif self.pool_mode == 'Max':
            o = mlp(mx)
        else:
            o = mlp(avg) + mlp(mx)
        s = torch.sigmoid(o).to(x.dtype).view(x.shape[0], x.shape[1], 1, 1)
        return s * x


class SLAM(nn.Module):
    """Spatial local attention (reference model/sradsgan.py:129-151)."""

    def __init__(self, kernel_size=7, pool_mode='Avg|Max'):
        super().__init__()
        assert kernel_size in (3, 7), 'kernel size must be 3 or 7'
        padding = 3 if kernel_size == 7 else 1
        self.pool_mode = pool_mode
        self.conv1 = Conv2d(2 if pool_mode == 'Avg|Max' else 1, 1, kernel_size, padding=padding, bias=False)

    def forward(self, x):
        x = ops.to_compute(x)
        if self.pool_mode == 'Avg':
            q = x.mean(dim=1, keepdim=True, dtype=torch.float32)
        elif self.pool_mode == 'Max':
            q = torch.max(x, dim=1, keepdim=True)[0].float()
        else:
            q = torch.cat([x.mean(dim=1, keepdim=True, dtype=torch.float32),
                           torch.max(x, dim=1, keepdim=True)[0].float()], dim=1)
        m = torch.sigmoid(self.conv1(q).float()).to(x.dtype)
        return m * x


class SGAM(nn.Module):
    """Position (spatial) global attention (reference model/sradsgan.py:153-176). Logits/softmax in fp32."""

    def __init__(self, in_dim):
        super().__init__()
        self.chanel_in = in_dim
        self.query_conv = Conv2d(in_dim, in_dim // 8, 1)
        self.key_conv = Conv2d(in_dim, in_dim // 8, 1)
        self.value_conv = Conv2d(in_dim, in_dim, 1)
        self.gamma = Parameter(torch.zeros(1))

    def forward(self, x):
        b, c, h, w = x.shape
        if (ops.config.compute_dtype == torch.bfloat16 and x.is_cuda and c == 64 and self.query_conv.out_channels == 8
                and h * w >= 64):
            # flash-style kernels (csrc/sgam.cu): no (HW) x (HW) energy / softmax tensors
            return ops.sgam_attention(self.query_conv(x), self.key_conv(x), self.value_conv(x), x, self.gamma)
        xf = x.float()
        q = self.query_conv(x).float().flatten(2).permute(0, 2, 1)
        k = self.key_conv(x).float().flatten(2)
        att = torch.softmax(torch.bmm(q, k), dim=-1)
        v = self.value_conv(x).float().flatten(2)
        out = torch.bmm(v, att.permute(0, 2, 1)).view(b, c, h, w)
        return self.gamma * out + xf


class CGAM(nn.Module):
    """Channel global attention (reference model/sradsgan.py:178-213): softmax(rowmax(XX^T) - XX^T) X. fp32."""

    def __init__(self, in_dim, light=False):
        super().__init__()
        self.chanel_in = in_dim
        self.light = light
        if light:
            self.conv1x1 = Conv2d(in_dim * 2, in_dim, 1, 1, bias=True)
        self.gamma = Parameter(torch.zeros(1))

    def forward(self, x):
        b, c, h, w = x.shape
        if not self.light and c == 64:
            return ops.cgam_attention(x, self.gamma)         # gram / softmax / apply kernels (csrc/cgam.cu), fp32
        xf = x.float()
        if self.light:
            pooled = torch.cat([F.adaptive_avg_pool2d(xf, 1), F.adaptive_max_pool2d(xf, 1)], 1)
            p = torch.relu(self.conv1x1(pooled).float()).view(b, c, -1)
            energy = torch.bmm(p, p.permute(0, 2, 1))
        else:
            q = xf.flatten(2)
            energy = torch.bmm(q, q.permute(0, 2, 1))
        energy_new = torch.max(energy, -1, keepdim=True)[0].expand_as(energy) - energy
        att = torch.softmax(energy_new, dim=-1)
        out = torch.bmm(att, xf.flatten(2)).view(b, c, h, w)
        return self.gamma * out + xf


def _la_init(mod, planes, la_mode, pool_mode, addconv):
    mod.la_mode, mod.addconv = la_mode, addconv
    if la_mode.find('CA') != -1:
        mod.ca = CLAM(planes, pool_mode=pool_mode)
    if la_mode.find('SA') != -1:
        mod.sa = SLAM(kernel_size=7, pool_mode=pool_mode)
    if la_mode.find('|') != -1:
        mod.conv = Conv2d(planes * 2, planes, 1, bias=True)
    if la_mode.find('-') != -1 and addconv:
        mod.conv = Conv2d(planes, planes, 1, bias=True)
    if la_mode == '':
        mod.last_conv = Conv2d(64, 64, 1, bias=True)


def _fused_la(mod):
    """the default local-attention configuration, the one the fused chain kernels implement"""
    return (mod.la_mode == 'CA-SA' and mod.addconv and mod.ca.pool_mode == 'Avg|Max' and mod.sa.pool_mode == 'Avg|Max'
            and mod.sa.conv1.kernel_size == 7 and mod.ca.fc1.out_channels <= 16)


def _la_forward(mod, out, x, acc=None, want_pool=False):
    """local-attention tail shared by RAB (:254-274) and ResGroup (:303-323), ending with `out += x`
    (fused into the closing 1x1 conv as a residual epilogue when there is one).  The residual stream `x`
    and the block output stay in fp32: only the branch runs in the compute dtype, so rounding errors do not
    compound along the 36-block trunk (SURVEY.md "bf16 error budget")."""
    m = mod.la_mode
    x = x.float()
    f32 = torch.float32
    if m == 'CA':
        return mod.ca(out).float() + x
    if m == 'SA':
        return mod.sa(out).float() + x
    if out.shape[1] == 64 and _fused_la(mod):
        return ops.local_attn_chain(out, x, mod.ca, mod.sa, mod.conv, acc=acc, want_pool=want_pool)      # the default configuration: one fused chain
    if m in ('CA-SA', 'SA-CA'):
        out = mod.sa(mod.ca(out)) if m == 'CA-SA' else mod.ca(mod.sa(out))
        return mod.conv.fused(out, residual=x, out_dtype=f32) if mod.addconv else out.float() + x
    if m == 'CA|SA':
        return mod.conv.fused(torch.cat([mod.ca(out), mod.sa(out)], dim=1), residual=x, out_dtype=f32)
    if m == '':
        return mod.last_conv.fused(out, residual=x, out_dtype=f32)
    return out.float() + x


class RAB(nn.Module):
    """Residual attention block (reference model/sradsgan.py:215-275)."""

    def __init__(self, inplanes, planes, kernel_size=3, stride=1, padding=1, bias=True, dilation=1, act_type='lrelu',
                 la_mode='CA-SA', pool_mode='Avg|Max', addconv=True):
        super().__init__()
        self.inplanes, self.planes = inplanes, planes
        self.conv1 = Conv2d(inplanes, 4 * planes, kernel_size, stride, padding, bias=bias, dilation=dilation)
        self.conv2 = Conv2d(4 * planes, planes, kernel_size, stride, padding, bias=bias, dilation=dilation)
        _la_init(self, planes, la_mode, pool_mode, addconv)
        self.act_type = act_type
        if act_type == 'prelu':
            self.act = nn.PReLU()
        elif act_type in ('tanh', 'sigmoid'):
            self.act = nn.Tanh() if act_type == 'tanh' else nn.Sigmoid()

    def forward(self, x, want_pool=False):
        """want_pool (new, internal): also emit the CLAM pooling partials of the output for the ResGroup tail that reads it"""
        xc = ops.to_compute(x)
        c1, c2 = self.conv1, self.conv2
        if (self.act_type in ('lrelu', 'relu') and c1.kernel_size == 3 and c1.stride == 1 and c1.padding == 1 and c1.bias is not None
                and c2.bias is not None):
            # the default block: both convolutions in one autograd node (activation derivative fused into conv2's dgrad); conv2's
            # epilogue emits the pooling partials the chain's CLAM needs
            out = ops.conv_act_conv(xc, c1, c2, ACT_LRELU if self.act_type == 'lrelu' else ACT_RELU, 0.2, want_pool=_fused_la(self))
            return _la_forward(self, out, x, want_pool=want_pool)
        if self.act_type == 'lrelu':
            out = self.conv1.fused(xc, ACT_LRELU, 0.2)
        elif self.act_type == 'relu':
            out = self.conv1.fused(xc, ACT_RELU)
        elif self.act_type in ('prelu', 'tanh', 'sigmoid'):
            out = self.act(self.conv1.fused(xc).float()).to(xc.dtype)
        else:
            out = self.conv1.fused(xc)
        out = self.conv2.fused(out)
        return _la_forward(self, out, x)


class ResGroup(nn.Module):
    """Residual group (reference model/sradsgan.py:277-324); like the reference it always builds its
    blocks with act_type='lrelu' (:284-285)."""

    def __init__(self, block, n_blocks=5, nc=64, kernel_size=3, stride=1, bias=True, padding=1,
                 act_type='lrelu', mode='CNA', rla_mode='CA-SA', bla_mode='CA-SA', pool_mode='Avg|Max', addconv=True):
        super().__init__()
        self.RG = nn.Sequential(*[block(nc, nc, kernel_size=kernel_size, bias=bias, stride=stride, padding=padding,
                                        act_type='lrelu', la_mode=bla_mode, pool_mode=pool_mode, addconv=addconv)
                                  for _ in range(n_blocks)])
        _la_init(self, nc, rla_mode, pool_mode, addconv)

    def forward(self, x, acc=None):
        """acc (new, internal): the generator's dense-sampling accumulator; when given, the sum `out_all += y` (reference :459)
        is done in the epilogue of the chain kernel and rides along as `y._sr_acc` (the module still returns y alone, so
        forward hooks and callers see the reference's output)"""
        fused = _fused_la(self)
        t = x
        for i, blk in enumerate(self.RG):
            last = i == len(self.RG) - 1
            t = blk(t, want_pool=True) if (last and fused and isinstance(blk, RAB)) else blk(t)
        if acc is None:
            return _la_forward(self, ops.to_compute(t), x)
        if fused:
            y, new_acc = _la_forward(self, ops.to_compute(t), x, acc=acc)
        else:
            y = _la_forward(self, ops.to_compute(t), x)
            new_acc = acc + y
        y._sr_acc = new_acc
        return y


class MSB(nn.Module):
    """Multi-scale block (reference model/sradsgan.py:326-345)."""

    def __init__(self, inplanes, planes):
        super().__init__()
        self.inplanes, self.planes = inplanes, planes
        self.conv1 = Conv2d(inplanes, planes, 3, 1, 1)
        self.conv2 = nn.Sequential(Conv2d(inplanes, planes, 1, bias=True), Conv2d(planes, planes, 3, 1, 1))
        self.conv3 = Conv2d(inplanes, planes, 1, bias=True)
        self.conv = Conv2d(planes * 3, planes, 1, bias=True)
        self.lrelu = LeakyReLU(inplace=True)

    def forward(self, x):
        x = ops.to_compute(x)
        out1 = self.conv1.fused(x)
        out2 = self.conv2[1].fused(self.conv2[0].fused(x))
        out3 = self.conv3.fused(x)
        return self.conv.fused(torch.cat([out1, out2, out3], dim=1), ACT_LRELU, self.lrelu.negative_slope,
                               out_dtype=torch.float32)


class ACB(nn.Module):
    """Asymmetric conv block (reference model/sradsgan.py:347-363) — unused by the generator; the 1x3/3x1
    branches are expressed as zero-padded 3x3 kernels are NOT needed on the hot path, so it is omitted."""

    def __init__(self, inplanes, planes):
        super().__init__()
        raise NotImplementedError("ACB is dead code in the reference generator (model/sradsgan.py:444-445)")


class GAB_UP(nn.Module):
    """Global attention + weight-tied sub-pixel upsampler (reference model/sradsgan.py:365-418)."""

    def __init__(self, ga_mode='CA-SA', addconv=True, upscale_factor=4):
        super().__init__()
        self.ga_mode, self.addconv = ga_mode, addconv
        if ga_mode.find('CA') != -1:
            self.ca = CGAM(64)
        if ga_mode.find('SA') != -1:
            self.sa = SGAM(64)
        if ga_mode.find('-') != -1 and addconv:
            self.conv = Conv2d(64, 64, 1, bias=True)
        if ga_mode.find('|') != -1:
            self.conv = Conv2d(64 * 2, 64, 1, bias=True)
        upsampling = []
        two = [Conv2d(64, 64 * 4, 3, 1, 1), PixelShuffle(2), LeakyReLU()]
        three = [Conv2d(64, 64 * 9, 3, 1, 1), PixelShuffle(3), LeakyReLU()]
        if (upscale_factor & (upscale_factor - 1)) == 0:
            for _ in range(int(math.log(upscale_factor, 2))):
                upsampling += two          # the SAME modules re-appended: stages share one conv (:388-392)
        elif upscale_factor % 3 == 0:
            for _ in range(int(math.log(upscale_factor, 3))):
                upsampling += three
        self.upsampling = nn.Sequential(*upsampling)
        if len(upsampling) > 3:            # one conv applied at several stages: its gradient is a sum over the uses
            for p in upsampling[0].parameters():
                p._sr_shared = True

    def forward(self, x):
        out = x
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
            out = self.conv.fused(torch.cat([self.ca(out), self.sa(out)], dim=1))
        mods = list(self.upsampling)
        for i in range(0, len(mods), 3):   # conv -> PixelShuffle(r) -> LeakyReLU fused into one kernel
            out = mods[i].fused(out, ACT_LRELU, mods[i + 2].negative_slope, shuffle_r=mods[i + 1].upscale_factor)
        return out


class GeneratorResNet(nn.Module):
    """SRADSGAN generator (reference model/sradsgan.py:420-468): MSB + conv1, n residual groups whose
    outputs are SUMMED into a dense-sampling accumulator (kept in fp32), GAB_UP, 3x3 output conv (fp32)."""

    def __init__(self, buildingblock, in_channels=3, out_channels=3, n_residual_blocks=12, n_basic_blocks=3,
                 rla_mode='CA-SA', bla_mode='CA-SA', ga_mode='CA-SA', pool_mode='Avg|Max', addconv=True, upscale_factor=4):
        super().__init__()
        self.conv1 = nn.Sequential(Conv2d(in_channels, 64, 3, 1, 1), LeakyReLU(inplace=True))
        self.res_groups = nn.Sequential(*[
            buildingblock(RAB, n_blocks=n_basic_blocks, nc=64, kernel_size=3, stride=1, padding=1, act_type='lrelu',
                          mode='CNA', rla_mode=rla_mode, bla_mode=bla_mode, pool_mode=pool_mode, addconv=addconv)
            for _ in range(n_residual_blocks)])
        self.GAB_UP = GAB_UP(ga_mode=ga_mode, addconv=addconv, upscale_factor=upscale_factor)
        self.MSB = MSB(inplanes=in_channels, planes=64)
        self.conv3 = nn.Sequential(Conv2d(64, out_channels, 3, 1, 1))

    def forward(self, x):
        x = ops.to_compute(x)
        msb = self.MSB(x)
        out = self.conv1[0].fused(x, ACT_LRELU, self.conv1[1].negative_slope, out_dtype=torch.float32)
        out_all = msb.float() + out
        for res_group in self.res_groups:
            if isinstance(res_group, ResGroup):
                y = res_group(out, acc=out_all)               # fp32 residual stream; out_all += y inside the chain kernel
                out_all = y._sr_acc if getattr(y, "_sr_acc", None) is not None else out_all + y
            else:
                y = res_group(out)
                out_all = out_all + y
            out = y
        out_all = self.GAB_UP(out_all)
        return self.conv3[0].fused(out_all, out_dtype=torch.float32)


class ChannelAttention(CLAM):
    """reference model/base_networks.py:366-403 (used by Discriminator); same arithmetic as CLAM."""


class SpatialAttention(SLAM):
    """reference model/base_networks.py:424-457 (used by Discriminator); same arithmetic as SLAM."""


class Discriminator(nn.Module):
    """Strided-conv PatchGAN critic with BatchNorm and CBAM attention after block 6
    (reference model/sradsgan.py:470-508).  Every op is differentiable twice (WGAN-GP)."""

    def __init__(self, in_channels=3, attention=True):
        super().__init__()
        layers = []
        in_filters = in_channels
        for layer, out_filters, stride, normalize in [(1, 64, 1, False), (2, 64, 2, True), (3, 128, 1, True),
                                                      (4, 128, 2, True), (5, 256, 1, True), (6, 256, 2, True),
                                                      (7, 512, 1, True), (8, 512, 2, True)]:
            layers.append(Conv2d(in_filters, out_filters, 3, stride, 1))
            if normalize:
                layers.append(BatchNorm2d(out_filters))
            layers.append(LeakyReLU(0.2, inplace=True))
            if attention and layer == 6:
                layers.append(ChannelAttention(256))
                layers.append(SpatialAttention())
            in_filters = out_filters
        layers.append(Conv2d(out_filters, 1, 3, 1, 1))
        self.model = nn.Sequential(*layers)
        self.block_taps = None       # tests: dict filled with the output of every fused block, keyed by the index of its last module

    def forward(self, img):
        x = ops.to_compute(img)
        # Every pass — G-step critic, D real / fake, and the WGAN-GP pass that is differentiated twice (reference
        # :611-621,:639) — runs the same fused kernels: conv(+bias+LReLU) and train-mode BatchNorm+LReLU Functions
        # whose backward passes are themselves differentiable Functions.  `ops.config.double_backward` (tests) and
        # eval mode take the unfused module-by-module path.
        if ops.config.double_backward or not self.training:
            return self.model(x)
        taps = self.block_taps
        i, n = 0, len(self.model)
        counters = []                 # num_batches_tracked of the BatchNorm layers of this pass: advanced together below
        while i < n:
            x, i = self.run_block(i, x, counters)
            if taps is not None:
                taps[i - 1] = x.detach()
        if counters:
            with torch.no_grad():
                torch._foreach_add_(counters, 1)
        return x

    def run_block(self, i, x, counters=None):
        """one fused block of the training path starting at Sequential index i: conv -> BatchNorm(train) + LeakyReLU,
        conv + LeakyReLU (epilogue), or a single module.  Returns (output, index of the next block).  `counters`: list that
        receives the BatchNorm layer's `num_batches_tracked` instead of it being advanced here."""
        mods = self.model
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        nxt2 = mods[i + 2] if i + 2 < len(mods) else None
        if isinstance(m, Conv2d) and isinstance(nxt, BatchNorm2d) and isinstance(nxt2, LeakyReLU):
            if counters is not None:
                counters.append(nxt.num_batches_tracked)
            return ops.bn_leaky_relu(m(x), nxt, nxt2.negative_slope, bump=counters is None), i + 3             # conv -> fused BN+LReLU
        if isinstance(m, Conv2d) and isinstance(nxt, LeakyReLU):
            return ops.conv2d_act(x, m.weight, m.bias, m.stride, m.padding, ACT_LRELU, nxt.negative_slope), i + 2   # conv + LReLU epilogue
        if (isinstance(m, ChannelAttention) and isinstance(nxt, SpatialAttention) and m.pool_mode == 'Avg|Max' and nxt.pool_mode == 'Avg|Max'
                and x.shape[1] % 64 == 0 and ops.config.fused_d_attention):
            return ops.cbam_attention(x, m.fc1.weight, m.fc2.weight, nxt.conv1), i + 2                              # csrc/cbam.cu primitives
        return m(x), i + 1


from .trainer import SRADSGAN  # noqa: E402,F401  (reference: class SRADSGAN lives in this module, :510)
