"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's EDSR workload (SURVEY.md §8 f1,
BASELINE.json configs[4]).  Nothing under `sradsgan_b200/` may import this file.

Follows /root/reference/SRADSGAN/model/edsr.py (`Net` :23-75, one training iteration :246-265) and the two
blocks it uses from model/base_networks.py (`ConvBlock` :170-208, `ResnetBlock` :246-298, both with norm=None),
functionally over a state_dict with the reference's keys, on the same ATen primitives.

PINNING: tests/test_oracle_vs_reference.py (imports the unmodified `model.edsr.Net` through oracle/ref_shim.py,
<=1e-5) and tests/golden/edsr_golden.pt (made by oracle/make_golden_edsr.py from the imported reference).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .sradsgan_oracle import _conv, make_state, synthetic_batch, unique_params, upsample_stages  # noqa: F401


def edsr_spec(scale=4, n_res=32, nf=256, in_ch=3):
    """Key order/shape of Net.state_dict() (model/edsr.py:23-60); the up-sampling convs are hard-wired to 256 (:42-48)."""
    s = OrderedDict()
    _conv(s, "input_conv.conv", nf, in_ch, 3)                       # :27
    for i in range(n_res):
        _conv(s, "residual_layers.%d.conv1" % i, nf, nf, 3)         # ResnetBlock base_networks.py:249-250
        _conv(s, "residual_layers.%d.conv2" % i, nf, nf, 3)
    _conv(s, "mid_conv.conv", nf, nf, 3)                            # :34
    r, n = upsample_stages(scale)
    for i in range(n):                                              # the SAME Conv2d re-appended per stage (:50-55)
        _conv(s, "upsampling.%d" % (3 * i), 256 * r * r, 256, 3)
    _conv(s, "output_conv.conv", in_ch, nf, 3)                      # :60
    return s


def tie_upsampling(sd):
    """state_dict() of the reference lists the shared conv under every stage index; alias them to one tensor"""
    for k in list(sd.keys()):
        if k.startswith("upsampling.") and not k.startswith("upsampling.0."):
            sd[k] = sd["upsampling.0." + k.split(".", 2)[2]]
    return sd


def edsr_forward(sd, x, scale=4, n_res=32, taps=None):
    """Net.forward (model/edsr.py:66-75)"""
    def c(name, t):
        return F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=1)
    out = c("input_conv.conv", x)                                   # :67
    residual = out
    for i in range(n_res):                                          # ResnetBlock.forward base_networks.py:283-297
        h = F.relu(c("residual_layers.%d.conv1" % i, out))
        out = c("residual_layers.%d.conv2" % i, h) + out
        if taps is not None:
            taps["residual_layers.%d" % i] = out
    out = c("mid_conv.conv", out) + residual                        # :70-71
    r, n = upsample_stages(scale)
    for i in range(n):                                              # :73  conv -> PixelShuffle -> LeakyReLU(0.01)
        out = F.leaky_relu(F.pixel_shuffle(c("upsampling.%d" % (3 * i), out), r), 0.01)
        if taps is not None:
            taps["upsampling.%d" % (3 * i)] = out
    return c("output_conv.conv", out)                               # :74


class EdsrTrainState:
    def __init__(self, sd, scale=4, n_res=32, lr=1e-4, b1=0.9, b2=0.999):
        self.sd, self.scale, self.n_res = sd, scale, n_res
        for p in unique_params(sd):
            p.requires_grad_(True)
        self.opt = torch.optim.Adam(unique_params(sd), lr=lr, betas=(b1, b2))     # :184


def edsr_train_step(st, imgs_lr, imgs_hr):
    """one iteration of EDSR.train (model/edsr.py:252-265): L1 pixel loss, backward, Adam"""
    st.opt.zero_grad()
    gen_hr = edsr_forward(st.sd, imgs_lr, st.scale, st.n_res)
    loss = F.l1_loss(gen_hr, imgs_hr)
    loss.backward()
    st.opt.step()
    return {"loss_G": loss.item(), "gen_hr": gen_hr.detach()}
