"""TEST INFRASTRUCTURE ONLY — generates tests/golden/ndsrgan_golden.pt from the UNMODIFIED reference `model.ndsrgan` classes.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_ndsrgan
The full 23-block generator on small inputs (outputs + gradient summaries), and two full training iterations driven exactly as
model/ndsrgan.py:414-456 does around the imported `GeneratorResNet` / `Discriminator` (Smooth-L1 criteria, Adam).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ndsrgan_oracle as N  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle import sradsgan_oracle as O  # noqa: E402
from oracle.make_golden import GOLDEN_DIR, build_ref_vgg, summarize  # noqa: E402

# name, scale, batch, LR size
NDSRGAN_CASES = [("ndsrgan_x4", 4, 2, 8), ("ndsrgan_x3", 3, 2, 6), ("ndsrgan_x2", 2, 1, 8)]
TRAIN_CFG = dict(scale=4, batch=2, lr_size=8, gseed=81, dseed=82, vseed=83, data_seed=91, steps=2, lr=2e-4)


def build_g(ref, sd, scale):
    net = ref.GeneratorResNet(in_channels=3, out_channels=3, nf=64, nc=32, upscale_factor=scale)
    assert list(net.state_dict().keys()) == list(sd.keys()), "NDSRGAN generator key mismatch"
    net.load_state_dict(sd, strict=True)
    return net.train()


def main():
    ref = ref_shim.load_reference("model.ndsrgan")
    out = {}
    for name, scale, batch, lrs in NDSRGAN_CASES:
        wseed, dseed = 40 + scale, 70 + scale
        sd = N.make_gen_state(scale, 23, wseed)
        net = build_g(ref, sd, scale)
        lr, hr = N.synthetic_batch(batch, scale, lrs * scale, seed=dseed)
        y = net(lr)
        loss = torch.nn.SmoothL1Loss()(y, hr)
        loss.backward()
        out[name] = {"cfg": dict(scale=scale, batch=batch, lr_size=lrs, wseed=wseed, dseed=dseed), "out": y.detach().clone(),
                     "loss": loss.item(),
                     "grads": {k: summarize(p.grad, 8) for k, p in net.named_parameters()} if name == "ndsrgan_x4" else {}}
        print(name, tuple(y.shape), loss.item(), y.abs().max().item())
    c = TRAIN_CFG
    gsd = N.make_gen_state(c["scale"], 23, c["gseed"])
    dsd = N.make_state(N.discriminator_spec(), seed=c["dseed"], init="fan")
    vsd = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    G = build_g(ref, gsd, c["scale"])
    D = ref.Discriminator()
    assert list(D.state_dict().keys()) == list(dsd.keys()), "NDSRGAN discriminator key mismatch"
    D.load_state_dict(dsd, strict=True)
    D.train()
    V = build_ref_vgg(vsd)
    for p in V.parameters():
        p.requires_grad_(False)
    opt_G = torch.optim.Adam(G.parameters(), lr=c["lr"], betas=(0.9, 0.999))      # model/ndsrgan.py:348-349
    opt_D = torch.optim.Adam(D.parameters(), lr=c["lr"], betas=(0.9, 0.999))
    crit = torch.nn.SmoothL1Loss()                                                  # :325-329
    steps = []
    for it in range(c["steps"]):
        lr, hr = N.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        opt_G.zero_grad()                                                           # :414
        gen_hr = G(lr)
        validity = D(gen_hr)
        loss_gan = crit(validity, torch.ones_like(validity))                        # :420
        loss_content = crit(V(gen_hr), V(hr).detach())                              # :423-425
        pix = crit(gen_hr, hr)                                                      # :429
        loss_G = 1e-2 * pix + loss_content + 2.5e-3 * loss_gan                      # :432
        loss_G.backward()
        opt_G.step()
        opt_D.zero_grad()                                                           # :441
        d_real, d_fake = D(hr), D(gen_hr.detach())
        loss_D = (crit(d_real, torch.ones_like(d_real)) + crit(d_fake, torch.zeros_like(d_fake))) / 2   # :444-451
        loss_D.backward()
        opt_D.step()
        steps.append({"loss_G": loss_G.item(), "loss_D": loss_D.item(), "pixel": pix.item(), "content": loss_content.item(),
                      "adv": loss_gan.item(),
                      "G": {k: summarize(v.float(), 8) for k, v in G.state_dict().items()} if it == c["steps"] - 1 else {},
                      "D": {k: summarize(v.float(), 8) for k, v in D.state_dict().items()}})
        print("step", it, loss_G.item(), loss_D.item())
    out["train_steps"] = {"cfg": c, "steps": steps}
    path = os.path.join(GOLDEN_DIR, "ndsrgan_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
