"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's NDSRGAN sibling.
Nothing under `sradsgan_b200/` may import this file.

Follows /root/reference/SRADSGAN/model/ndsrgan.py: `CL` :57-58, `DenseBlock` :60-77, `DCRDB` :79-93, `DRRDBnet` :95-169,
`GeneratorResNet` :171-223, `Discriminator` :225-258 and one training iteration :414-456 (every criterion torch.nn.SmoothL1Loss,
:325-329), functionally over state_dicts with the reference's keys, on the same ATen primitives.

PINNING: tests/test_ndsrgan_cpu.py (imports the unmodified `model.ndsrgan` classes through oracle/ref_shim.py, <=1e-5) and
tests/golden/ndsrgan_golden.pt (made by oracle/make_golden_ndsrgan.py from the imported reference).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .sradsgan_oracle import _c, _conv, make_state, synthetic_batch, unique_params, upsample_stages, vgg_features, vgg_spec  # noqa: F401
from .srgan_oracle import _batch_norm, _bn, noise_grad_keys  # noqa: F401

D_BLOCKS = [(64, 2, False), (128, 2, True), (256, 2, True), (512, 1, True)]       # model/ndsrgan.py:240-243


def generator_spec(scale=4, n_blocks=23, nf=64, nc=32, in_ch=3, out_ch=3):
    s = OrderedDict()
    _conv(s, "conv1.0", nf, in_ch, 3)
    for k in range(1, n_blocks + 1):
        p = "DCRDB_block.DRRDB%d" % k
        for r in (1, 2, 3):
            for j in range(4):
                _conv(s, "%s.RDB%d.CL_blocks.%d.0" % (p, r, j), nc, nc * j + nf, 3)
            _conv(s, "%s.RDB%d.conv" % (p, r), nf, nc * 4 + nf, 3)
        _conv(s, p + ".conv", nf, nf, 3)
    _conv(s, "conv2", nf, nf, 3)
    r, n = upsample_stages(scale)
    for i in range(n):
        _conv(s, "upsampling.%d" % (3 * i + 1), nf, nf, 3)
    _conv(s, "conv3.0", nf, nf, 3)
    _conv(s, "conv3.2", out_ch, nf, 3)
    return s


def tie_upsampling(sd):
    for k in list(sd.keys()):
        if k.startswith("upsampling."):
            idx, rest = k.split(".", 2)[1:]
            if int(idx) >= 3:
                sd[k] = sd["upsampling.%d.%s" % ((int(idx) - 1) % 3 + 1, rest)]
    return sd


def discriminator_spec(in_ch=3):
    s, idx, cin = OrderedDict(), 0, in_ch
    for cout, stride, norm in D_BLOCKS:
        _conv(s, "model.%d" % idx, cout, cin, 4); idx += 1
        if norm:
            _bn(s, "model.%d" % idx, cout); idx += 1
        idx += 1
        cin = cout
    _conv(s, "model.%d" % idx, 1, cin, 4)
    return s


def dense_block(sd, p, x):
    out1 = x
    for j in range(4):
        y = F.leaky_relu(_c(sd, "%s.CL_blocks.%d.0" % (p, j), x), 0.2)
        x = torch.cat((x, y), dim=1)
    return out1 + _c(sd, p + ".conv", x) * 0.2


def dcrdb(sd, p, x):
    out1 = dense_block(sd, p + ".RDB1", x)
    out2 = dense_block(sd, p + ".RDB2", x + 0.2 * out1)
    out3 = dense_block(sd, p + ".RDB3", x + 0.2 * out1 + 0.2 * out2)
    out4 = _c(sd, p + ".conv", x + 0.2 * out1 + 0.2 * out2 + 0.2 * out3)
    return out4 * 0.2 + x


def generator_forward(sd, x, scale=4, n_blocks=23, taps=None):
    out = _c(sd, "conv1.0", x)
    acc = out
    for k in range(1, n_blocks + 1):                     # m_k = DRRDB_k(x + 0.2 m_1 + ... + 0.2 m_{k-1}), summed left to right (:121-168)
        m = dcrdb(sd, "DCRDB_block.DRRDB%d" % k, acc)
        if taps is not None:
            taps["DRRDB%d" % k] = m
        acc = acc + 0.2 * m
    out = out + _c(sd, "conv2", acc)
    r, n = upsample_stages(scale)
    for i in range(n):
        out = F.leaky_relu(_c(sd, "upsampling.%d" % (3 * i + 1), F.interpolate(out, scale_factor=r, mode="nearest")), 0.2)
    return _c(sd, "conv3.2", F.leaky_relu(_c(sd, "conv3.0", out), 0.2))


def discriminator_forward(sd, img, update_stats=True):
    x, idx = img, 0
    for cout, stride, norm in D_BLOCKS:
        x = F.conv2d(x, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=stride, padding=1); idx += 1
        if norm:
            x = _batch_norm(sd, "model.%d" % idx, x, update_stats); idx += 1
        x = F.leaky_relu(x, 0.2); idx += 1
    return F.conv2d(x, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=1, padding=1)


class TrainState:
    def __init__(self, G, D, V, scale=4, n_blocks=23, lr=2e-4, b1=0.9, b2=0.999):
        self.G, self.D, self.V, self.scale, self.n_blocks = G, D, V, scale, n_blocks
        for p in unique_params(G) + unique_params(D):
            p.requires_grad_(True)
        for p in unique_params(V):
            p.requires_grad_(False)
        self.opt_G = torch.optim.Adam(unique_params(G), lr=lr, betas=(b1, b2))
        self.opt_D = torch.optim.Adam(unique_params(D), lr=lr, betas=(b1, b2))


def train_step(st, imgs_lr, imgs_hr):
    """one iteration of NDSRGAN.train (model/ndsrgan.py:414-456)"""
    sl1 = F.smooth_l1_loss
    st.opt_G.zero_grad()
    gen_hr = generator_forward(st.G, imgs_lr, st.scale, st.n_blocks)
    validity = discriminator_forward(st.D, gen_hr)
    loss_gan = sl1(validity, torch.ones_like(validity))
    content = sl1(vgg_features(st.V, gen_hr), vgg_features(st.V, imgs_hr).detach())
    pix = sl1(gen_hr, imgs_hr)
    loss_G = 1e-2 * pix + content + 2.5e-3 * loss_gan
    loss_G.backward()
    st.opt_G.step()
    st.opt_D.zero_grad()
    d_real, d_fake = discriminator_forward(st.D, imgs_hr), discriminator_forward(st.D, gen_hr.detach())
    loss_D = (sl1(d_real, torch.ones_like(d_real)) + sl1(d_fake, torch.zeros_like(d_fake))) / 2
    loss_D.backward()
    st.opt_D.step()
    return {"loss_G": loss_G.item(), "loss_D": loss_D.item(), "pixel": pix.item(), "content": content.item(), "adv": loss_gan.item(),
            "gen_hr": gen_hr.detach()}


def make_gen_state(scale, n_blocks, seed, gain=0.55):
    """fan-in initialised generator weights with a gain below one: the 0.2-scaled dense skips of 23 x 3 blocks otherwise grow the
    activations to O(100), which says nothing about parity"""
    sd = make_state(generator_spec(scale, n_blocks), seed=seed, init="fan")
    for k, v in sd.items():
        if v.dim() == 4:
            v.mul_(gain / 1.3)
    return tie_upsampling(sd)
