"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the SRADSGAN hot path (the parity oracle).

Nothing under `sradsgan_b200/` may import this file; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do (as the checker / the CPU baseline).

The reference is pure Python over torch ATen ops (torch==1.8.1 pinned in `requirements.txt:2`, not
vendored); this file restates its algorithm functionally over a `state_dict` (same keys/shapes as the
reference modules) using the same ATen primitives on CPU in fp32.  Every function cites the reference
lines it follows (paths relative to /root/reference/SRADSGAN).

PINNING: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the reference's own classes run in the build container:
  * tests/test_oracle_vs_reference.py  — imports the reference through oracle/ref_shim.py (skipped when
    /root/reference is absent) and checks forward, losses, gradients and a full G+D step to <=1e-5;
  * tests/golden/*.pt (made by oracle/make_golden.py from the imported reference) — checked everywhere.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# state_dict specifications (key -> shape), mirroring the reference constructors
# ----------------------------------------------------------------------------------------------


def _conv(spec, name, cout, cin, k, bias=True):
    spec[name + ".weight"] = (cout, cin, k, k)
    if bias:
        spec[name + ".bias"] = (cout,)


def _la(spec, prefix, nc=64, ratio=16):
    # CLAM model/sradsgan.py:110-112 (no bias), SLAM :136 (2->1, 7x7, no bias), 1x1 conv :233/:297
    _conv(spec, prefix + ".ca.fc1", nc // ratio, nc, 1, bias=False)
    _conv(spec, prefix + ".ca.fc2", nc, nc // ratio, 1, bias=False)
    _conv(spec, prefix + ".sa.conv1", 1, 2, 7, bias=False)
    _conv(spec, prefix + ".conv", nc, nc, 1)


def upsample_stages(scale):
    """model/sradsgan.py:387-392 -> (r, number of stages)."""
    if (scale & (scale - 1)) == 0:
        return 2, int(math.log(scale, 2))
    if scale % 3 == 0:
        return 3, int(math.log(scale, 3))
    return 1, 0


def generator_spec(scale=4, n_groups=12, n_blocks=3, in_ch=3, out_ch=3):
    """Key order/shape of GeneratorResNet.state_dict() (model/sradsgan.py:420-448)."""
    s = OrderedDict()
    _conv(s, "conv1.0", 64, in_ch, 3)
    for g in range(n_groups):
        for b in range(n_blocks):
            p = "res_groups.%d.RG.%d" % (g, b)
            _conv(s, p + ".conv1", 256, 64, 3)      # RAB :222
            _conv(s, p + ".conv2", 64, 256, 3)      # RAB :223
            _la(s, p)
        _la(s, "res_groups.%d" % g)                 # ResGroup :290-297
    s["GAB_UP.ca.gamma"] = (1,)                     # CGAM :187
    s["GAB_UP.sa.gamma"] = (1,)                     # SGAM :162
    _conv(s, "GAB_UP.sa.query_conv", 8, 64, 1)
    _conv(s, "GAB_UP.sa.key_conv", 8, 64, 1)
    _conv(s, "GAB_UP.sa.value_conv", 64, 64, 1)
    _conv(s, "GAB_UP.conv", 64, 64, 1)
    r, n = upsample_stages(scale)
    for i in range(n):                              # tied: every stage is the SAME Conv2d (:381-392)
        _conv(s, "GAB_UP.upsampling.%d" % (3 * i), 64 * r * r, 64, 3)
    _conv(s, "MSB.conv1", 64, in_ch, 3)
    _conv(s, "MSB.conv2.0", 64, in_ch, 1)
    _conv(s, "MSB.conv2.1", 64, 64, 3)
    _conv(s, "MSB.conv3", 64, in_ch, 1)
    _conv(s, "MSB.conv", 64, 192, 1)
    _conv(s, "conv3.0", out_ch, 64, 3)
    return s


D_BLOCKS = [(64, 1, False), (64, 2, True), (128, 1, True), (128, 2, True),
            (256, 1, True), (256, 2, True), (512, 1, True), (512, 2, True)]


def discriminator_layout(in_ch=3):
    """Sequential index layout of Discriminator.model (model/sradsgan.py:482-505).
    Returns a list of ('conv', idx, cin, cout, stride) / ('bn', idx, c) / ('lrelu',) / ('ca', idx, c) /
    ('sa', idx)."""
    layers, idx, cin = [], 0, in_ch
    for li, (cout, stride, norm) in enumerate(D_BLOCKS, start=1):
        layers.append(("conv", idx, cin, cout, stride)); idx += 1
        if norm:
            layers.append(("bn", idx, cout)); idx += 1
        layers.append(("lrelu",)); idx += 1
        if li == 6:                                   # :494-496 (the `layers == 8` branch :497 is dead)
            layers.append(("ca", idx, 256)); idx += 1
            layers.append(("sa", idx)); idx += 1
        cin = cout
    layers.append(("conv", idx, cin, 1, 1))
    return layers


def discriminator_spec(in_ch=3):
    s = OrderedDict()
    for l in discriminator_layout(in_ch):
        if l[0] == "conv":
            _conv(s, "model.%d" % l[1], l[3], l[2], 3)
        elif l[0] == "bn":
            p = "model.%d" % l[1]
            s[p + ".weight"] = (l[2],); s[p + ".bias"] = (l[2],)
            s[p + ".running_mean"] = (l[2],); s[p + ".running_var"] = (l[2],)
            s[p + ".num_batches_tracked"] = ()
        elif l[0] == "ca":                            # base_networks.py:380-382, ratio 16
            _conv(s, "model.%d.fc1" % l[1], l[2] // 16, l[2], 1, bias=False)
            _conv(s, "model.%d.fc2" % l[1], l[2], l[2] // 16, 1, bias=False)
        elif l[0] == "sa":                            # base_networks.py:436
            _conv(s, "model.%d.conv1" % l[1], 1, 2, 7, bias=False)
    return s


VGG_CFG = [(0, 3, 64), (2, 64, 64), "M", (5, 64, 128), (7, 128, 128), "M", (10, 128, 256)]


def vgg_spec():
    """torchvision vgg19().features[:12] as used by FeatureExtractor (model/sradsgan.py:92-95);
    keys carry the `feature_extractor.` prefix of the reference module."""
    s = OrderedDict()
    for c in VGG_CFG:
        if c != "M":
            _conv(s, "feature_extractor.%d" % c[0], c[2], c[1], 3)
    return s


# ----------------------------------------------------------------------------------------------
# seeded synthetic weights
# ----------------------------------------------------------------------------------------------

def make_state(spec, seed, init="ref", gamma=0.5):
    """Deterministic synthetic weights.

    init="ref": utils/utils.py:97-114 (`weights_init_normal` as applied by train(), sradsgan.py:713-714):
        conv W ~ N(0,0.02), b = 0, BatchNorm gamma ~ N(1,0.02), beta = 0.
    init="fan": W ~ N(0, 1/fan_in) * 1.3, b ~ N(0, 0.05) — O(1) activations, for well-conditioned
        relative-error checks (the 0.02 init yields ~1e-3 outputs).
    CGAM/SGAM gamma (untouched by the reference init, = 0) are set to `gamma` so the attention is exercised.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for k, shp in spec.items():
        if k.endswith("gamma"):
            sd[k] = torch.full(shp, float(gamma))
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            sd[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(shp)
        elif len(shp) == 4:
            if init == "ref":
                sd[k] = torch.randn(shp, generator=g) * 0.02
            else:
                fan_in = shp[1] * shp[2] * shp[3]
                sd[k] = torch.randn(shp, generator=g) * (1.3 / math.sqrt(fan_in))
        else:  # 1-D: conv bias or BN weight/bias
            parent_is_bn = (k.rsplit(".", 1)[0] + ".running_mean") in spec
            if parent_is_bn and k.endswith(".weight"):
                sd[k] = 1.0 + torch.randn(shp, generator=g) * 0.02
            elif init == "ref" or parent_is_bn:
                sd[k] = torch.zeros(shp)
            else:
                sd[k] = torch.randn(shp, generator=g) * 0.05
    return sd


def tie_upsampling(sd):
    """The reference re-appends one Conv2d object (model/sradsgan.py:388-392): aliases share storage."""
    for i in (3, 6):
        for t in ("weight", "bias"):
            k = "GAB_UP.upsampling.%d.%s" % (i, t)
            if k in sd:
                sd[k] = sd["GAB_UP.upsampling.0." + t]
    return sd


def synthetic_batch(batch, scale=4, hr_size=216, seed=1234):
    """hr ~ U[0,1), lr = bicubic(hr) clamped to [0,1]."""
    g = torch.Generator().manual_seed(seed)
    hr = torch.rand(batch, 3, hr_size, hr_size, generator=g)
    lr = F.interpolate(hr, size=hr_size // scale, mode="bicubic", align_corners=False).clamp(0, 1)
    return lr, hr


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------

def _c(sd, name, x, stride=1, pad=None):
    w = sd[name + ".weight"]
    b = sd.get(name + ".bias")
    if pad is None:
        pad = w.shape[-1] // 2
    return F.conv2d(x, w, b, stride=stride, padding=pad)


def clam(sd, p, x):
    """CLAM.forward, pool_mode 'Avg|Max' (model/sradsgan.py:117-127); ChannelAttention in D
    (base_networks.py:387-403) is the same computation."""
    w1, w2 = sd[p + ".fc1.weight"], sd[p + ".fc2.weight"]
    avg = F.adaptive_avg_pool2d(x, 1)
    mx = F.adaptive_max_pool2d(x, 1)
    a = F.conv2d(F.relu(F.conv2d(avg, w1)), w2)
    m = F.conv2d(F.relu(F.conv2d(mx, w1)), w2)
    return torch.sigmoid(a + m) * x


def slam(sd, p, x):
    """SLAM.forward (model/sradsgan.py:141-151); SpatialAttention in D (base_networks.py:441-457)."""
    avg = torch.mean(x, dim=1, keepdim=True)
    mx, _ = torch.max(x, dim=1, keepdim=True)
    q = torch.cat([avg, mx], dim=1)
    return torch.sigmoid(F.conv2d(q, sd[p + ".conv1.weight"], padding=3)) * x


def la_chain(sd, p, x):
    """'CA-SA' + addconv branch shared by RAB (:258-262) and ResGroup (:307-311)."""
    x = clam(sd, p + ".ca", x)
    x = slam(sd, p + ".sa", x)
    return _c(sd, p + ".conv", x)


def rab(sd, p, x, taps=None):
    """RAB.forward (model/sradsgan.py:250-275): conv3x3 64->256, LeakyReLU(0.2), conv3x3 256->64,
    CLAM, SLAM, conv1x1, += x."""
    out = _c(sd, p + ".conv1", x)
    out = F.leaky_relu(out, 0.2)
    out = _c(sd, p + ".conv2", out)
    if taps is not None:
        taps[p + ".conv2"] = out
    out = la_chain(sd, p, out)
    return out + x


def res_group(sd, p, x, n_blocks, taps=None):
    """ResGroup.forward (model/sradsgan.py:301-324)."""
    out = x
    for b in range(n_blocks):
        out = rab(sd, "%s.RG.%d" % (p, b), out, taps)
        if taps is not None:
            taps["%s.RG.%d" % (p, b)] = out
    out = la_chain(sd, p, out)
    return out + x


def msb(sd, x):
    """MSB.forward (model/sradsgan.py:339-345); LeakyReLU default slope 0.01 (:337)."""
    o1 = _c(sd, "MSB.conv1", x)
    o2 = _c(sd, "MSB.conv2.1", _c(sd, "MSB.conv2.0", x))
    o3 = _c(sd, "MSB.conv3", x)
    return F.leaky_relu(_c(sd, "MSB.conv", torch.cat([o1, o2, o3], dim=1)), 0.01)


def cgam(sd, p, x):
    """CGAM.forward, light=False (model/sradsgan.py:202-213): softmax(rowmax(E) - E), E = X X^T."""
    b, c, h, w = x.shape
    q = x.reshape(b, c, -1)
    energy = torch.bmm(q, q.permute(0, 2, 1))
    energy_new = torch.max(energy, -1, keepdim=True)[0].expand_as(energy) - energy
    att = torch.softmax(energy_new, dim=-1)
    out = torch.bmm(att, q).reshape(b, c, h, w)
    return sd[p + ".gamma"] * out + x


def sgam(sd, p, x):
    """SGAM.forward (model/sradsgan.py:164-176): position attention, q,k in R^8, v in R^64."""
    b, c, h, w = x.shape
    q = _c(sd, p + ".query_conv", x).reshape(b, -1, h * w).permute(0, 2, 1)
    k = _c(sd, p + ".key_conv", x).reshape(b, -1, h * w)
    att = torch.softmax(torch.bmm(q, k), dim=-1)
    v = _c(sd, p + ".value_conv", x).reshape(b, -1, h * w)
    out = torch.bmm(v, att.permute(0, 2, 1)).reshape(b, c, h, w)
    return sd[p + ".gamma"] * out + x


def gab_up(sd, x, scale, taps=None):
    """GAB_UP.forward, ga_mode 'CA-SA' (model/sradsgan.py:396-418) with the weight-tied upsampler."""
    out = cgam(sd, "GAB_UP.ca", x)
    if taps is not None:
        taps["GAB_UP.ca"] = out
    out = sgam(sd, "GAB_UP.sa", out)
    if taps is not None:
        taps["GAB_UP.sa"] = out
    out = _c(sd, "GAB_UP.conv", out)
    r, n = upsample_stages(scale)
    for i in range(n):
        out = _c(sd, "GAB_UP.upsampling.0", out)
        out = F.leaky_relu(F.pixel_shuffle(out, r), 0.01)
        if taps is not None:
            taps["GAB_UP.up_stage.%d" % i] = out
    return out


def generator_forward(sd, x, scale=4, n_groups=12, n_blocks=3, taps=None):
    """GeneratorResNet.forward (model/sradsgan.py:450-468). Dense sampling is a running SUM."""
    m = msb(sd, x)
    out = F.leaky_relu(_c(sd, "conv1.0", x), 0.01)
    if taps is not None:
        taps["MSB"] = m
        taps["conv1"] = out
    out_all = m + out
    for g in range(n_groups):
        y = res_group(sd, "res_groups.%d" % g, out, n_blocks, taps)
        if taps is not None:
            taps["res_groups.%d" % g] = y
        out_all = out_all + y
        out = y
    if taps is not None:
        taps["out_all"] = out_all
    up = gab_up(sd, out_all, scale, taps)
    return _c(sd, "conv3.0", up)


def discriminator_forward(sd, img, update_stats=True, taps=None):
    """Discriminator.forward in train mode (model/sradsgan.py:470-508): BatchNorm2d uses batch
    statistics and (when update_stats) updates running_mean/var in `sd` (momentum .1, eps 1e-5)."""
    x = img
    for l in discriminator_layout(img.shape[1]):
        if l[0] == "conv":
            x = _c(sd, "model.%d" % l[1], x, stride=l[4])
        elif l[0] == "bn":
            p = "model.%d" % l[1]
            rm = sd[p + ".running_mean"] if update_stats else None
            rv = sd[p + ".running_var"] if update_stats else None
            x = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], training=True,
                             momentum=0.1, eps=1e-5)
            if update_stats:
                sd[p + ".num_batches_tracked"] += 1
        elif l[0] == "lrelu":
            x = F.leaky_relu(x, 0.2)
        elif l[0] == "ca":
            x = clam(sd, "model.%d" % l[1], x)
        elif l[0] == "sa":
            x = slam(sd, "model.%d" % l[1], x)
        if taps is not None and l[0] != "lrelu":
            taps["model.%d" % l[1]] = x
    return x


def vgg_features(sd, img):
    """FeatureExtractor.forward (model/sradsgan.py:97-99): vgg19.features[:12], raw [0,1] input."""
    x = img
    for c in VGG_CFG:
        if c == "M":
            x = F.max_pool2d(x, 2, 2)
        else:
            x = F.relu(_c(sd, "feature_extractor.%d" % c[0], x))
    return x


def wgan_loss(pred, target_is_real):
    """GANLoss('wgan-gp') (model/sradsgan.py:46-52)."""
    return -pred.mean() if target_is_real else pred.mean()


def psnr(pred, gt):
    """utils/utils.py:700-709."""
    mse = torch.mean((pred.clamp(0, 1) - gt.clamp(0, 1)).double() ** 2).item()
    return 100.0 if mse == 0 else 10 * math.log10(1.0 / mse)


def quantize_u8(img):
    """utils/utils.py:169-175 (`save_img1`): *255, clamp, astype(uint8) == truncation. CHW -> HWC."""
    return (img * 255.0).clamp(0, 255).detach().numpy().transpose(1, 2, 0).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# one training iteration (model/sradsgan.py:829-892 + :595-641)
# ----------------------------------------------------------------------------------------------

# Parameters whose exact gradient is identically ZERO, so the reference's Adam turns fp rounding noise
# into +-lr steps (not reproducible by any other summation order; excluded from post-step parity):
#   * SGAM key bias: adds q_i.b to every logit of row i -> softmax shift-invariant (model/sradsgan.py:167-169)
#   * every D conv bias that feeds a train-mode BatchNorm2d (model/sradsgan.py:476-478)
NOISE_GRAD_KEYS = ("GAB_UP.sa.key_conv.bias",) + tuple("model.%d.bias" % i for i in (2, 5, 8, 11, 14, 19, 22))


def unique_params(sd):
    """Parameters as nn.Module.parameters() yields them: unique tensors, buffers excluded."""
    seen, out = set(), []
    for k, v in sd.items():
        if k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            continue
        if id(v) in seen:
            continue
        seen.add(id(v))
        out.append(v)
    return out


def gradient_penalty(D, real, fake, alpha, norm="L2", penalty="LS"):
    """SRADSGAN.gradient_penalty (model/sradsgan.py:595-641). `alpha` (B,1,1,1) is the caller's draw
    of np.random.random (:609). Calls .backward() itself (:639) and returns the scalar."""
    inter = (alpha * real + (1 - alpha) * fake).requires_grad_(True)
    d = discriminator_forward(D, inter)
    grads = torch.autograd.grad(outputs=d, inputs=inter, grad_outputs=torch.ones_like(d),
                                create_graph=True, retain_graph=True, only_inputs=True)[0]
    if norm == "Linf":
        gn, _ = torch.max(torch.abs(grads), 1)
    elif norm == "L1":
        gn = grads.norm(1, 1)
    else:
        gn = grads.norm(2, 1)                       # per-pixel norm over the 3 channels (:630)
    c = (gn - 1).pow(2) if penalty == "LS" else torch.relu(gn - 1)
    gp = c.mean()
    gp.backward(retain_graph=True)
    return gp


class TrainState:
    """Leaf parameters + Adam optimisers for one (G, D, VGG) triple."""

    def __init__(self, G, D, V, scale=4, n_groups=12, n_blocks=3, lr=2e-4, b1=0.9, b2=0.999,
                 weight_content=1e-2, weight_gan=1e-3, lambda_gp=10.0, clip_value=0.01):
        self.G, self.D, self.V = G, D, V
        self.scale, self.n_groups, self.n_blocks = scale, n_groups, n_blocks
        self.wc, self.wg, self.lgp, self.clip = weight_content, weight_gan, lambda_gp, clip_value
        for p in unique_params(G) + unique_params(D):
            p.requires_grad_(True)
        for p in unique_params(V):
            p.requires_grad_(False)   # VGG grads are computed-and-discarded in the reference
        self.opt_G = torch.optim.Adam(unique_params(G), lr=lr, betas=(b1, b2))   # :724
        self.opt_D = torch.optim.Adam(unique_params(D), lr=lr, betas=(b1, b2))   # :725


def train_step(st, imgs_lr, imgs_hr, alpha):
    """One iteration of SRADSGAN.train (model/sradsgan.py:829-892). Returns dict of scalars + gen_hr."""
    st.opt_G.zero_grad()
    gen_hr = generator_forward(st.G, imgs_lr, st.scale, st.n_groups, st.n_blocks)       # :832
    pixel = F.l1_loss(gen_hr, imgs_hr)                                                   # :834
    gen_f = vgg_features(st.V, gen_hr)                                                   # :836
    real_f = vgg_features(st.V, imgs_hr).detach()                                        # :837
    content = F.l1_loss(gen_f, real_f)                                                   # :838
    adv = wgan_loss(discriminator_forward(st.D, gen_hr), True)                           # :847-848
    loss_G = pixel + st.wc * content + st.wg * adv                                       # :852
    loss_G.backward()
    st.opt_G.step()                                                                      # :857-858
    st.opt_D.zero_grad()                                                                 # :865
    loss_real = wgan_loss(discriminator_forward(st.D, imgs_hr), True)                    # :876
    loss_fake = wgan_loss(discriminator_forward(st.D, gen_hr.detach()), False)           # :877
    loss_D = loss_real + loss_fake
    gp = gradient_penalty(st.D, imgs_hr.detach(), gen_hr.detach(), alpha)                # :882 (backward #1)
    loss_D = loss_D + st.lgp * gp                                                        # :884
    loss_D.backward()                                                                    # :886 (backward #2)
    st.opt_D.step()
    with torch.no_grad():
        for p in unique_params(st.D):
            p.clamp_(-st.clip, st.clip)                                                  # :891-892
    return {"loss_G": loss_G.item(), "loss_D": loss_D.item(), "pixel": pixel.item(),
            "content": content.item(), "adv": adv.item(), "gp": gp.item(), "gen_hr": gen_hr.detach()}
