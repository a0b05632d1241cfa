"""TEST INFRASTRUCTURE ONLY — generates tests/golden/sragan_golden.pt from the UNMODIFIED reference `model.sragan` classes.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_sragan
Generator outputs / gradient summaries for several scales, and two full training iterations driven exactly as
model/sragan.py:642-705 does around the imported `GeneratorResNet`, `Discriminator`, `GANLoss` and `SRAGAN.gradient_penalty`
(oracle.make_golden.ref_train_step: the iteration is SRADSGAN's line for line).  Weights are regenerated from the seeds by the tests.
"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402
from oracle import sradsgan_oracle as O  # noqa: E402
from oracle import sragan_oracle as A  # noqa: E402
from oracle.make_golden import GOLDEN_DIR, build_ref_vgg, ref_train_step, summarize  # noqa: E402

# name, scale, residual blocks, basic blocks, batch, LR size
SRAGAN_CASES = [("sragan_x4", 4, 2, 2, 2, 12), ("sragan_x2", 2, 1, 3, 2, 10), ("sragan_x3", 3, 1, 1, 2, 9), ("sragan_x9", 9, 1, 2, 2, 5)]
TRAIN_CFG = dict(scale=4, n_res=2, n_basic=2, batch=2, lr_size=8, gseed=51, dseed=52, vseed=53, data_seed=61, np_seed=71, steps=2)


def build_g(ref, sd, scale, n_res, n_basic):
    net = ref.GeneratorResNet(ref.ResidualBlock_Block_WithAttention, n_residual_blocks=n_res, n_basic_blocks=n_basic, rla_mode='CA-SA',
                              bla_mode='CA-SA', ga_mode='CA-SA', pool_mode='Avg|Max', upscale_factor=scale)
    assert list(net.state_dict().keys()) == list(sd.keys()), "SRAGAN generator key mismatch"
    net.load_state_dict(sd, strict=True)
    return net.train()


def main():
    ref = ref_shim.load_reference("model.sragan")
    out = {}
    for name, scale, n_res, n_basic, batch, lrs in SRAGAN_CASES:
        wseed, dseed = 30 + scale, 90 + scale
        sd = A.tie_upsampling(A.make_state(A.generator_spec(scale, n_res, n_basic), seed=wseed, init="fan"))
        net = build_g(ref, sd, scale, n_res, n_basic)
        lr, hr = A.synthetic_batch(batch, scale, lrs * scale, seed=dseed)
        y = net(lr)
        loss = torch.nn.L1Loss()(y, hr)
        loss.backward()
        out[name] = {"cfg": dict(scale=scale, n_res=n_res, n_basic=n_basic, batch=batch, lr_size=lrs, wseed=wseed, dseed=dseed),
                     "out": y.detach().clone(), "loss": loss.item(),
                     "grads": {k: summarize(p.grad, 8) for k, p in net.named_parameters()}}
        print(name, tuple(y.shape), loss.item())
    c = TRAIN_CFG
    gsd = A.tie_upsampling(A.make_state(A.generator_spec(c["scale"], c["n_res"], c["n_basic"]), seed=c["gseed"], init="fan"))
    dsd = O.make_state(O.discriminator_spec(), seed=c["dseed"], init="ref")
    vsd = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    G = build_g(ref, gsd, c["scale"], c["n_res"], c["n_basic"])
    D = ref.Discriminator()
    assert list(D.state_dict().keys()) == list(dsd.keys()), "SRAGAN discriminator key mismatch"
    D.load_state_dict(dsd, strict=True)
    V = build_ref_vgg(vsd)
    shim = types.SimpleNamespace(GANLoss=ref.GANLoss, SRADSGAN=ref.SRAGAN)       # ref_train_step calls <module>.SRADSGAN.gradient_penalty
    opt_G = torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.9, 0.999))      # model/sragan.py:537-538
    opt_D = torch.optim.Adam(D.parameters(), lr=2e-4, betas=(0.9, 0.999))
    steps = []
    for it in range(c["steps"]):
        lr, hr = A.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        rec = ref_train_step(shim, G, D, V, opt_G, opt_D, lr, hr, np_seed=c["np_seed"] + it)
        rec["G_state"] = {k: summarize(p.float(), 8) for k, p in G.state_dict().items()}
        rec["D_state"] = {k: summarize(p.float(), 8) for k, p in D.state_dict().items()}
        steps.append(rec)
        print("step", it, {k: v for k, v in rec.items() if isinstance(v, float)})
    out["train_steps"] = {"cfg": c, "steps": steps}
    path = os.path.join(GOLDEN_DIR, "sragan_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
