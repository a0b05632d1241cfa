"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt from the UNMODIFIED reference classes.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference's nets (`GeneratorResNet`, `Discriminator`, `GANLoss`, and `SRADSGAN.gradient_penalty`
called unbound) are imported through oracle/ref_shim.py and fed the seeded weights/inputs of
`oracle/sradsgan_oracle.make_state` / `synthetic_batch`.  `train()` itself needs datasets, TF1 and CUDA
tensors, so one iteration is driven here exactly as model/sradsgan.py:829-892 does, around the imported
modules.  Weights are NOT stored (they are regenerated from the seed); only outputs/summaries are.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402
from oracle import sradsgan_oracle as O  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def summarize(t, n=48):
    """norm / mean / n pseudo-randomly sampled values (fixed LCG indices) of a tensor."""
    f = t.detach().double().reshape(-1)
    idx = (torch.arange(n, dtype=torch.long) * 2654435761 + 12345) % f.numel()
    return {"shape": tuple(t.shape), "norm": f.norm().item(), "mean": f.mean().item(),
            "samples": f[idx].float().clone()}


def build_ref_generator(ref, sd, scale, n_groups, n_blocks):
    g = ref.GeneratorResNet(ref.ResGroup, n_residual_blocks=n_groups, n_basic_blocks=n_blocks,
                            rla_mode='CA-SA', bla_mode='CA-SA', ga_mode='CA-SA', pool_mode='Avg|Max',
                            upscale_factor=scale)
    assert list(g.state_dict().keys()) == list(sd.keys()), "generator key mismatch"
    g.load_state_dict(sd, strict=True)
    return g


def build_ref_discriminator(ref, sd):
    d = ref.Discriminator()
    assert list(d.state_dict().keys()) == list(sd.keys()), "discriminator key mismatch"
    d.load_state_dict(sd, strict=True)
    return d


class RefFeatureExtractor(nn.Module):
    """FeatureExtractor (model/sradsgan.py:88-99) minus the pretrained download (no network)."""

    def __init__(self):
        super().__init__()
        from torchvision.models import vgg19
        self.feature_extractor = nn.Sequential(*list(vgg19(weights=None).features.children())[:12])

    def forward(self, img):
        return self.feature_extractor(img)


def build_ref_vgg(sd):
    v = RefFeatureExtractor()
    assert list(v.state_dict().keys()) == list(sd.keys()), "vgg key mismatch"
    v.load_state_dict(sd, strict=True)
    return v


def ref_train_step(ref, G, D, V, opt_G, opt_D, imgs_lr, imgs_hr, np_seed,
                   weight_content=1e-2, weight_gan=1e-3, lambda_gp=10.0, clip_value=0.01):
    """model/sradsgan.py:829-892 around the imported reference modules (non-relativistic branch)."""
    crit = torch.nn.L1Loss()
    gan = ref.GANLoss(gan_type='wgan-gp', real_label_val=1.0, fake_label_val=0.0)
    fake_self = types.SimpleNamespace(gpu_mode=False)
    rec = {}
    opt_G.zero_grad()
    gen_hr = G(imgs_lr)
    pixel = crit(gen_hr, imgs_hr)
    gen_f = V(gen_hr)
    real_f = V(imgs_hr).data
    content = crit(gen_f, real_f)
    adv = gan(D(gen_hr), True)
    loss_G = pixel + weight_content * content + weight_gan * adv
    loss_G.backward()
    rec["G_grads"] = {k: summarize(p.grad, 8) for k, p in G.named_parameters()}
    opt_G.step()
    opt_D.zero_grad()
    loss_real = gan(D(imgs_hr), True)
    loss_fake = gan(D(gen_hr.detach()), False)
    loss_D = loss_real + loss_fake
    np.random.seed(np_seed)
    gp = ref.SRADSGAN.gradient_penalty(fake_self, D, imgs_hr.data, gen_hr.detach().data,
                                       grad_penalty_Lp_norm='L2', penalty_type='LS')
    loss_D += lambda_gp * gp
    loss_D.backward()
    rec["D_grads"] = {k: summarize(p.grad, 8) for k, p in D.named_parameters()}
    opt_D.step()
    for p in D.parameters():
        p.data.clamp_(-clip_value, clip_value)
    rec.update({"loss_G": loss_G.item(), "loss_D": loss_D.item(), "pixel": pixel.item(),
                "content": content.item(), "adv": adv.item(), "gp": gp.item(),
                "gen_hr": summarize(gen_hr, 64)})
    return rec


GEN_CASES = [
    # name, scale, n_groups, n_blocks, batch, lr_size, init
    ("g_x4_small", 4, 2, 1, 2, 12, "fan"),
    ("g_x2_small", 2, 1, 2, 1, 10, "fan"),
    ("g_x3_small", 3, 1, 1, 1, 9, "fan"),
    ("g_x8_small", 8, 1, 1, 1, 6, "fan"),
    ("g_x9_small", 9, 1, 1, 1, 5, "fan"),
    ("g_x4_refinit", 4, 2, 1, 1, 12, "ref"),
    ("g_x4_full", 4, 12, 3, 1, 16, "fan"),
    ("g_x4_full_refinit", 4, 12, 3, 1, 16, "ref"),     # the init train() really applies (utils.py:97-114)
]


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_shim.load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    out = {}
    for name, scale, ng, nb, batch, lrs, init in GEN_CASES:
        sd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=100 + scale, init=init))
        G = build_ref_generator(ref, sd, scale, ng, nb)
        lr, hr = O.synthetic_batch(batch, scale, lrs * scale, seed=7 + scale)
        taps = {}
        hooks = []
        for mname, mod in G.named_modules():
            if mname and mname.count(".") <= 3 and not mname.endswith(("avg_pool", "max_pool", "relu1", "sigmoid",
                                                                         "softmax", "act", "lrelu")):
                hooks.append(mod.register_forward_hook(
                    lambda m, i, o, mname=mname: taps.__setitem__(mname, summarize(o, 16))))
        with torch.no_grad():
            y = G(lr)
        for h in hooks:
            h.remove()
        out[name] = {"cfg": dict(scale=scale, n_groups=ng, n_blocks=nb, batch=batch, lr_size=lrs, init=init,
                                 wseed=100 + scale, dseed=7 + scale),
                     "out": y.clone() if y.numel() <= 20000 else None, "out_sum": summarize(y, 256),
                     "psnr_vs_hr": O.psnr(y, hr), "taps": taps}
        print(name, tuple(y.shape), "psnr", out[name]["psnr_vs_hr"], "taps", len(taps))

    # discriminator forward + BN running-stat update
    dsd = O.make_state(O.discriminator_spec(), seed=11, init="fan")
    D = build_ref_discriminator(ref, {k: v.clone() for k, v in dsd.items()})
    x = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    dt = {}
    hooks = [mod.register_forward_hook(lambda m, i, o, n=n: dt.__setitem__(n, summarize(o, 16)))
             for n, mod in D.model.named_children()]
    y = D(x)
    for h in hooks:
        h.remove()
    out["d_fwd"] = {"out": y.detach().clone(), "taps": {"model." + k: v for k, v in dt.items()},
                    "bn": {k: v.clone() for k, v in D.state_dict().items() if "running" in k or "tracked" in k}}
    print("d_fwd", tuple(y.shape))

    # VGG19[:12]
    vsd = O.make_state(O.vgg_spec(), seed=12, init="fan")
    V = build_ref_vgg(vsd)
    xv = torch.rand(1, 3, 16, 16, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        out["vgg"] = {"out": V(xv).clone()}

    # two full training iterations, small generator
    scale, ng, nb, batch, lrs = 4, 2, 1, 2, 8
    gsd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=21, init="fan"))
    dsd = O.make_state(O.discriminator_spec(), seed=22, init="ref")
    vsd = O.make_state(O.vgg_spec(), seed=23, init="fan")
    G = build_ref_generator(ref, gsd, scale, ng, nb)
    D = build_ref_discriminator(ref, dsd)
    V = build_ref_vgg(vsd)
    opt_G = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.9, 0.999))
    opt_D = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.9, 0.999))
    steps = []
    for it in range(2):
        lr, hr = O.synthetic_batch(batch, scale, lrs * scale, seed=31 + it)
        rec = ref_train_step(ref, G, D, V, opt_G, opt_D, lr, hr, np_seed=41 + it)
        rec["G_params"] = {k: summarize(p, 8) for k, p in G.named_parameters()}
        rec["D_state"] = {k: summarize(p.float(), 8) for k, p in D.state_dict().items()}
        steps.append(rec)
        print("step", it, {k: v for k, v in rec.items() if isinstance(v, float)})
    out["train_steps"] = {"cfg": dict(scale=scale, n_groups=ng, n_blocks=nb, batch=batch, lr_size=lrs,
                                      gseed=21, dseed=22, vseed=23, data_seed=31, np_seed=41),
                          "steps": steps}
    path = os.path.join(GOLDEN_DIR, "sradsgan_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


FULLSIZE = dict(scale=4, batch=16, hr=216, gseed=0, dseed=1, vseed=2, data_seed=1234, np_seed=4242, gamma=0.5)


def fullsize():
    """ONE iteration of the UNMODIFIED reference at the benchmarked configuration (BASELINE.json configs[1]: full 12x3
    generator, discriminator, VGG19[:12], batch 16, LR 54^2 / HR 216^2; weights per SURVEY.md §8d) -> losses, output /
    gradient / parameter summaries in tests/golden/sradsgan_fullsize_golden.pt.  tests/test_oracle_golden.py holds the oracle
    to it on CPU, tests/test_gpu_fullsize_parity.py holds `graphed_step` to the oracle AND to these numbers on the GPU."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_shim.load_reference()
    c = FULLSIZE
    gsd = O.tie_upsampling(O.make_state(O.generator_spec(c["scale"]), seed=c["gseed"], init="ref", gamma=c["gamma"]))
    dsd = O.make_state(O.discriminator_spec(), seed=c["dseed"], init="ref")
    vsd = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    G = build_ref_generator(ref, gsd, c["scale"], 12, 3)
    D = build_ref_discriminator(ref, dsd)
    V = build_ref_vgg(vsd)
    opt_G = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.9, 0.999))
    opt_D = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.9, 0.999))
    lr, hr = O.synthetic_batch(c["batch"], c["scale"], c["hr"], seed=c["data_seed"])
    with torch.no_grad():
        y0 = G(lr)
    rec = ref_train_step(ref, G, D, V, opt_G, opt_D, lr, hr, np_seed=c["np_seed"])
    rec["psnr_vs_hr"] = O.psnr(y0, hr)
    rec["G_params"] = {k: summarize(p, 8) for k, p in G.named_parameters()}
    rec["D_state"] = {k: summarize(p.float(), 8) for k, p in D.state_dict().items()}
    out = {"cfg": c, "step": rec}
    path = os.path.join(GOLDEN_DIR, "sradsgan_fullsize_golden.pt")
    torch.save(out, path)
    print("fullsize", {k: v for k, v in rec.items() if isinstance(v, float)})
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if "--fullsize" in sys.argv:
        fullsize()
    else:
        main()
