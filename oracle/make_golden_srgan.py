"""TEST INFRASTRUCTURE ONLY — generates tests/golden/srgan_golden.pt from the UNMODIFIED reference `model.srgan` classes.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_srgan
The reference's trainer (`SRGAN.train`, model/srgan.py:247-520) needs datasets, TF1 logging, LPIPS weights and CUDA tensors, so the
iteration is driven here exactly as :343-381 does around the imported `GeneratorResNet` / `Discriminator`; the VGG19[:12]
extractor is built with the same seeded weights as the SRADSGAN golden (no pretrained checkpoint offline).
Weights are regenerated from the seeds by the tests; only outputs / summaries are stored.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402
from oracle import sradsgan_oracle as O  # noqa: E402
from oracle import srgan_oracle as S  # noqa: E402
from oracle.make_golden import GOLDEN_DIR, summarize  # noqa: E402

# name, scale, residual blocks, batch, LR size
SRGAN_CASES = [("srgan_x4", 4, 2, 2, 12), ("srgan_x2", 2, 1, 2, 10), ("srgan_x3", 3, 1, 2, 9), ("srgan_x9", 9, 1, 2, 5)]
TRAIN_CFG = dict(scale=4, n_res=2, batch=2, lr_size=8, gseed=31, dseed=32, vseed=33, data_seed=41, steps=2, lr=2e-4)


def build_g(ref, sd, scale, n_res):
    net = ref.GeneratorResNet(in_channels=3, out_channels=3, n_residual_blocks=n_res, upscale_factor=scale)
    assert list(net.state_dict().keys()) == list(sd.keys()), "SRGAN generator key mismatch"
    net.load_state_dict(sd, strict=True)
    return net.train()


def build_d(ref, sd):
    net = ref.Discriminator()
    assert list(net.state_dict().keys()) == list(sd.keys()), "SRGAN discriminator key mismatch"
    net.load_state_dict(sd, strict=True)
    return net.train()


class _Vgg(torch.nn.Module):
    """vgg19.features[:12] with the seeded weights of oracle.vgg_spec (torchvision's constructor, no download)"""

    def __init__(self, vsd):
        super().__init__()
        from torchvision.models import vgg19
        self.feature_extractor = torch.nn.Sequential(*list(vgg19(weights=None).features.children())[:12])
        self.load_state_dict(vsd, strict=True)

    def forward(self, x):
        return self.feature_extractor(x)


def main():
    ref = ref_shim.load_reference("model.srgan")
    out = {}
    for name, scale, n_res, batch, lrs in SRGAN_CASES:
        wseed, dseed = 20 + scale, 60 + scale
        sd = S.tie_upsampling(S.make_state(S.generator_spec(scale, n_res), seed=wseed, init="fan"))
        net = build_g(ref, sd, scale, n_res)
        lr, hr = S.synthetic_batch(batch, scale, lrs * scale, seed=dseed)
        y = net(lr)
        loss = torch.nn.MSELoss()(y, hr)
        loss.backward()
        out[name] = {"cfg": dict(scale=scale, n_res=n_res, batch=batch, lr_size=lrs, wseed=wseed, dseed=dseed),
                     "out": y.detach().clone(), "loss": loss.item(),
                     "grads": {k: summarize(p.grad, 8) for k, p in net.named_parameters()},
                     "buffers": {k: summarize(v.float(), 8) for k, v in net.state_dict().items() if "running" in k}}
        print(name, tuple(y.shape), loss.item())
    c = TRAIN_CFG
    gsd = S.tie_upsampling(S.make_state(S.generator_spec(c["scale"], c["n_res"]), seed=c["gseed"], init="fan"))
    dsd = S.make_state(S.discriminator_spec(), seed=c["dseed"], init="fan")
    vsd = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    G, D, V = build_g(ref, gsd, c["scale"], c["n_res"]), build_d(ref, dsd), _Vgg(vsd)
    for p in V.parameters():
        p.requires_grad_(False)
    opt_G = torch.optim.Adam(G.parameters(), lr=c["lr"], betas=(0.9, 0.999))      # model/srgan.py:274
    opt_D = torch.optim.Adam(D.parameters(), lr=c["lr"], betas=(0.9, 0.999))      # :275
    mse = torch.nn.MSELoss()                                                        # :261-263
    steps = []
    for it in range(c["steps"]):
        lr, hr = S.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        opt_G.zero_grad()                                                           # :343
        gen_hr = G(lr)                                                              # :346
        validity = D(gen_hr)                                                        # :348
        loss_gan = mse(validity, torch.ones_like(validity))                         # :349
        loss_content = mse(V(gen_hr), V(hr).detach())                               # :352-354
        mse_g = mse(gen_hr, hr)                                                     # :358
        loss_G = mse_g + 6e-3 * loss_content + 1e-3 * loss_gan                      # :361
        loss_G.backward()
        opt_G.step()
        opt_D.zero_grad()                                                           # :367
        d_real, d_fake = D(hr), D(gen_hr.detach())
        loss_D = (mse(d_real, torch.ones_like(d_real)) + mse(d_fake, torch.zeros_like(d_fake))) / 2     # :370-377
        loss_D.backward()
        opt_D.step()
        steps.append({"loss_G": loss_G.item(), "loss_D": loss_D.item(), "pixel": mse_g.item(), "content": loss_content.item(),
                      "adv": loss_gan.item(),
                      "G": {k: summarize(v.float(), 8) for k, v in G.state_dict().items()},
                      "D": {k: summarize(v.float(), 8) for k, v in D.state_dict().items()}})
        print("step", it, loss_G.item(), loss_D.item())
    out["train_steps"] = {"cfg": c, "steps": steps}
    path = os.path.join(GOLDEN_DIR, "srgan_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
