"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's SRAGAN sibling.
Nothing under `sradsgan_b200/` may import this file.

Follows /root/reference/SRADSGAN/model/sragan.py (`GeneratorResNet` :147-237, one training iteration :642-705 — identical to
SRADSGAN's, model/sradsgan.py:829-892) and model/base_networks.py (`BasicBlock` :958-1070, `ResidualBlock_Block_WithAttention`
:1505-1594, `ChannelAttention` / `SpatialAttention` :366-457, `CAM_Module` / `PAM_Module` :480-554), functionally over state_dicts
with the reference's keys, on the same ATen primitives.  The critic, VGG extractor, WGAN loss and gradient penalty are those of
oracle/sradsgan_oracle.py (the reference's classes are the same code).

PINNING: tests/test_sragan_cpu.py (imports the unmodified `model.sragan` classes through oracle/ref_shim.py, <=1e-5) and
tests/golden/sragan_golden.pt (made by oracle/make_golden_sragan.py from the imported reference).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from . import sradsgan_oracle as O
from .sradsgan_oracle import _c, _conv, _la, make_state, synthetic_batch, unique_params, upsample_stages  # noqa: F401
from .srgan_oracle import _batch_norm, _bn, noise_grad_keys  # noqa: F401


def _block(s, p):
    """BasicBlock: two ConvBlocks (their Conv2d sits under `.conv`), CA / SA / 1x1 conv"""
    _conv(s, p + ".conv1.conv", 64, 64, 3)
    _conv(s, p + ".conv2.conv", 64, 64, 3)
    _la(s, p)


def generator_spec(scale=4, n_res=12, n_basic=5, in_ch=3, out_ch=3):
    """Key order/shape of GeneratorResNet(ResidualBlock_Block_WithAttention, ...).state_dict() (model/sragan.py:147-211)"""
    s = OrderedDict()
    _conv(s, "conv1.0", 64, in_ch, 3)
    for i in range(n_res):
        p = "res_blocks.%d" % i
        for j in range(n_basic - 1):
            _block(s, "%s.blocks.%d" % (p, j))
        _block(s, p + ".last_conv")
        _la(s, p)
    _conv(s, "conv2.0", 64, 64, 3); _bn(s, "conv2.1", 64)
    s["ca.gamma"] = (1,)
    s["sa.gamma"] = (1,)
    _conv(s, "sa.query_conv", 8, 64, 1); _conv(s, "sa.key_conv", 8, 64, 1); _conv(s, "sa.value_conv", 64, 64, 1)
    _conv(s, "conv", 64, 64, 1)
    r, n = upsample_stages(scale)
    for i in range(n):
        _conv(s, "upsampling.%d" % (4 * i), 64 * r * r, 64, 3)
        _bn(s, "upsampling.%d" % (4 * i + 1), 64 * r * r)
    _conv(s, "conv3.0", out_ch, 64, 3)
    return s


def tie_upsampling(sd):
    for k in list(sd.keys()):
        if k.startswith("upsampling."):
            idx, rest = k.split(".", 2)[1:]
            if int(idx) >= 4:
                sd[k] = sd["upsampling.%d.%s" % (int(idx) % 4, rest)]
    return sd


def basic_block(sd, p, x, act):
    """BasicBlock.forward (base_networks.py:1019-1070), inplanes == planes, la_mode 'CA-SA' + addconv"""
    out = _c(sd, p + ".conv1.conv", x)
    if act:
        out = F.leaky_relu(out, 0.2)                     # ConvBlock(activation='lrelu') (base_networks.py:187-188)
    out = _c(sd, p + ".conv2.conv", out)
    out = O.la_chain(sd, p, out) + x
    return F.leaky_relu(out, 0.2) if act else out


def res_block(sd, p, x, n_basic, taps=None):
    """ResidualBlock_Block_WithAttention.forward (base_networks.py:1551-1594), mode 'CNA': the last block has no activation"""
    out = x
    for j in range(n_basic - 1):
        out = basic_block(sd, "%s.blocks.%d" % (p, j), out, True)
    out = basic_block(sd, p + ".last_conv", out, False)
    return O.la_chain(sd, p, out) + x


def generator_forward(sd, x, scale=4, n_res=12, n_basic=5, taps=None, update_stats=True):
    """GeneratorResNet.forward in train mode, ga_mode 'CA-SA' (model/sragan.py:213-237)"""
    out1 = F.leaky_relu(_c(sd, "conv1.0", x), 0.01)
    out = out1
    for i in range(n_res):
        out = res_block(sd, "res_blocks.%d" % i, out, n_basic)
        if taps is not None:
            taps["res_blocks.%d" % i] = out
    out2 = _batch_norm(sd, "conv2.1", _c(sd, "conv2.0", out), update_stats)
    out = out1 + out2
    out = O.cgam(sd, "ca", out)
    out = O.sgam(sd, "sa", out)
    out = _c(sd, "conv", out)
    if taps is not None:
        taps["ga"] = out
    r, n = upsample_stages(scale)
    for i in range(n):
        out = _c(sd, "upsampling.%d" % (4 * i), out)
        out = _batch_norm(sd, "upsampling.%d" % (4 * i + 1), out, update_stats)
        out = F.leaky_relu(F.pixel_shuffle(out, r), 0.01)
    return torch.tanh(_c(sd, "conv3.0", out))


class TrainState(O.TrainState):
    def __init__(self, G, D, V, scale=4, n_res=12, n_basic=5, **kw):
        super().__init__(G, D, V, scale=scale, **kw)
        self.n_res, self.n_basic = n_res, n_basic


def train_step(st, imgs_lr, imgs_hr, alpha):
    """one iteration of SRAGAN.train (model/sragan.py:642-705)"""
    st.opt_G.zero_grad()
    gen_hr = generator_forward(st.G, imgs_lr, st.scale, st.n_res, st.n_basic)            # :645
    pixel = F.l1_loss(gen_hr, imgs_hr)                                                   # :647
    content = F.l1_loss(O.vgg_features(st.V, gen_hr), O.vgg_features(st.V, imgs_hr).detach())   # :649-651
    adv = O.wgan_loss(O.discriminator_forward(st.D, gen_hr), True)                       # :660-661
    loss_G = pixel + st.wc * content + st.wg * adv                                       # :665
    loss_G.backward()
    st.opt_G.step()
    st.opt_D.zero_grad()
    loss_D = O.wgan_loss(O.discriminator_forward(st.D, imgs_hr), True) + O.wgan_loss(O.discriminator_forward(st.D, gen_hr.detach()), False)
    gp = O.gradient_penalty(st.D, imgs_hr.detach(), gen_hr.detach(), alpha)              # :695 (backward #1)
    loss_D = loss_D + st.lgp * gp
    loss_D.backward()                                                                    # :699 (backward #2)
    st.opt_D.step()
    with torch.no_grad():
        for p in unique_params(st.D):
            p.clamp_(-st.clip, st.clip)                                                  # :704-705
    return {"loss_G": loss_G.item(), "loss_D": loss_D.item(), "pixel": pixel.item(), "content": content.item(), "adv": adv.item(),
            "gp": gp.item(), "gen_hr": gen_hr.detach()}
