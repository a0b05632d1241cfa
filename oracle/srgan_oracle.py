"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's SRGAN sibling.
Nothing under `sradsgan_b200/` may import this file.

Follows /root/reference/SRADSGAN/model/srgan.py: `ResidualBlock` :57-71, `GeneratorResNet` :73-123, `Discriminator` :125-156
(the SRADSGAN critic WITHOUT the attention pair), `FeatureExtractor` :44-55 and one training iteration :343-381, functionally
over state_dicts with the reference's keys, on the same ATen primitives.

PINNING: tests/test_srgan_cpu.py (imports the unmodified `model.srgan` classes through oracle/ref_shim.py, <=1e-5) and
tests/golden/srgan_golden.pt (made by oracle/make_golden_srgan.py from the imported reference).
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from .sradsgan_oracle import D_BLOCKS, _conv, make_state, synthetic_batch, unique_params, upsample_stages, vgg_features, vgg_spec  # noqa: F401


def _bn(s, p, c):
    s[p + ".weight"] = (c,); s[p + ".bias"] = (c,)
    s[p + ".running_mean"] = (c,); s[p + ".running_var"] = (c,)
    s[p + ".num_batches_tracked"] = ()


def generator_spec(scale=4, n_res=16, in_ch=3, out_ch=3):
    """Key order/shape of GeneratorResNet.state_dict() (model/srgan.py:73-113).  The up-sampling stage list re-appends the SAME
    conv / BatchNorm objects per stage (:94-107), so x4 / x8 / x9 list them under every stage index (see tie_upsampling)."""
    s = OrderedDict()
    _conv(s, "conv1.0", 64, in_ch, 9)                                   # :78-82
    for i in range(n_res):                                              # ResidualBlock :61-65
        p = "res_blocks.%d.conv_block" % i
        _conv(s, p + ".0", 64, 64, 3); _bn(s, p + ".1", 64)
        _conv(s, p + ".3", 64, 64, 3); _bn(s, p + ".4", 64)
    _conv(s, "conv2.0", 64, 64, 3); _bn(s, "conv2.1", 64)               # :91
    r, n = upsample_stages(scale)
    for i in range(n):                                                  # conv, bn, shuffle, relu = 4 entries per stage
        _conv(s, "upsampling.%d" % (4 * i), 64 * r * r, 64, 3)
        _bn(s, "upsampling.%d" % (4 * i + 1), 64 * r * r)
    _conv(s, "conv3.0", out_ch, 64, 9)                                  # :112
    return s


def tie_upsampling(sd):
    for k in list(sd.keys()):
        if k.startswith("upsampling."):
            idx, rest = k.split(".", 2)[1:]
            if int(idx) >= 4:
                sd[k] = sd["upsampling.%d.%s" % (int(idx) % 4, rest)]
    return sd


def discriminator_spec(in_ch=3):
    """model/srgan.py:125-151: eight conv blocks (BatchNorm on all but the first) + the 512 -> 1 output conv, no attention"""
    s, idx, cin = OrderedDict(), 0, in_ch
    for cout, stride, norm in D_BLOCKS:
        _conv(s, "model.%d" % idx, cout, cin, 3); idx += 1
        if norm:
            _bn(s, "model.%d" % idx, cout); idx += 1
        idx += 1                                                        # LeakyReLU
        cin = cout
    _conv(s, "model.%d" % idx, 1, cin, 3)
    return s


def noise_grad_keys(spec):
    """conv biases directly in front of a BatchNorm: the batch mean removes them, so their gradient is identically zero and Adam
    turns the rounding noise of any implementation into +-lr steps — such entries cannot be compared between implementations"""
    out = set()
    for k in spec:
        if k.endswith(".bias") and "." in k[:-5] and not (k[:-5] + ".running_mean") in spec:
            head, idx = k[:-5].rsplit(".", 1)
            if idx.isdigit() and ("%s.%d.running_mean" % (head, int(idx) + 1)) in spec:
                out.add(k)
    return out


def _batch_norm(sd, p, x, update_stats=True):
    rm = sd[p + ".running_mean"] if update_stats else None
    rv = sd[p + ".running_var"] if update_stats else None
    y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], training=True, momentum=0.1, eps=1e-5)
    if update_stats:
        sd[p + ".num_batches_tracked"] += 1
    return y


def generator_forward(sd, x, scale=4, n_res=16, taps=None, update_stats=True):
    """GeneratorResNet.forward in train mode (model/srgan.py:115-123)"""
    c = lambda name, t, pad: F.conv2d(t, sd[name + ".weight"], sd[name + ".bias"], stride=1, padding=pad)
    out1 = F.relu(c("conv1.0", x, 4))                                   # :116
    out = out1
    for i in range(n_res):                                              # :70  x + conv_block(x)
        p = "res_blocks.%d.conv_block" % i
        h = F.relu(_batch_norm(sd, p + ".1", c(p + ".0", out, 1), update_stats))
        out = out + _batch_norm(sd, p + ".4", c(p + ".3", h, 1), update_stats)
        if taps is not None:
            taps["res_blocks.%d" % i] = out
    out2 = _batch_norm(sd, "conv2.1", c("conv2.0", out, 1), update_stats)   # :118
    out = out1 + out2                                                   # :119
    r, n = upsample_stages(scale)
    for i in range(n):                                                  # :120  conv -> BN -> PixelShuffle -> ReLU
        out = c("upsampling.%d" % (4 * i), out, 1)
        out = _batch_norm(sd, "upsampling.%d" % (4 * i + 1), out, update_stats)
        out = F.relu(F.pixel_shuffle(out, r))
        if taps is not None:
            taps["upsampling.%d" % (4 * i)] = out
    return torch.tanh(c("conv3.0", out, 4))                             # :121


def discriminator_forward(sd, img, update_stats=True):
    x, idx = img, 0
    for cout, stride, norm in D_BLOCKS:
        x = F.conv2d(x, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=stride, padding=1); idx += 1
        if norm:
            x = _batch_norm(sd, "model.%d" % idx, x, update_stats); idx += 1
        x = F.leaky_relu(x, 0.2); idx += 1
    return F.conv2d(x, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=1, padding=1)


class TrainState:
    def __init__(self, G, D, V, scale=4, n_res=16, lr=2e-4, b1=0.9, b2=0.999):
        self.G, self.D, self.V, self.scale, self.n_res = G, D, V, scale, n_res
        for p in unique_params(G) + unique_params(D):
            p.requires_grad_(True)
        for p in unique_params(V):
            p.requires_grad_(False)
        self.opt_G = torch.optim.Adam(unique_params(G), lr=lr, betas=(b1, b2))      # :274
        self.opt_D = torch.optim.Adam(unique_params(D), lr=lr, betas=(b1, b2))      # :275


def train_step(st, imgs_lr, imgs_hr):
    """one iteration of SRGAN.train (model/srgan.py:343-381)"""
    st.opt_G.zero_grad()
    gen_hr = generator_forward(st.G, imgs_lr, st.scale, st.n_res)                        # :346
    gen_validity = discriminator_forward(st.D, gen_hr)                                   # :348
    loss_gan = F.mse_loss(gen_validity, torch.ones_like(gen_validity))                   # :349 (valid = ones of the patch shape)
    content = F.mse_loss(vgg_features(st.V, gen_hr), vgg_features(st.V, imgs_hr).detach())   # :352-354
    mse = F.mse_loss(gen_hr, imgs_hr)                                                    # :358
    loss_G = mse + 6e-3 * content + 1e-3 * loss_gan                                      # :361
    loss_G.backward()
    st.opt_G.step()
    st.opt_D.zero_grad()
    d_real = discriminator_forward(st.D, imgs_hr)
    d_fake = discriminator_forward(st.D, gen_hr.detach())
    loss_D = (F.mse_loss(d_real, torch.ones_like(d_real)) + F.mse_loss(d_fake, torch.zeros_like(d_fake))) / 2   # :373-380
    loss_D.backward()
    st.opt_D.step()
    return {"loss_G": loss_G.item(), "loss_D": loss_D.item(), "pixel": mse.item(), "content": content.item(),
            "adv": loss_gan.item(), "gen_hr": gen_hr.detach()}
