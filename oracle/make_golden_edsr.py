"""TEST INFRASTRUCTURE ONLY — generates tests/golden/edsr_golden.pt from the UNMODIFIED reference `model.edsr.Net`.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_edsr
The reference's trainer (`EDSR.train`, model/edsr.py:150-390) needs datasets, TF1 logging and CUDA tensors, so the
iteration is driven here exactly as :252-265 does (zero_grad, forward, L1, backward, Adam step) around the imported
`Net`.  Weights are regenerated from the seed by the tests; only outputs / summaries are stored.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import edsr_oracle as E  # noqa: E402
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import GOLDEN_DIR, summarize  # noqa: E402

# name, scale, residual blocks, batch, LR size
EDSR_CASES = [("edsr_x4", 4, 2, 2, 12), ("edsr_x2", 2, 1, 1, 10), ("edsr_x3", 3, 1, 1, 9), ("edsr_x8", 8, 1, 1, 6)]
TRAIN_CFG = dict(scale=4, n_res=2, batch=2, lr_size=12, wseed=61, data_seed=71, steps=2, lr=1e-4)


def build(ref, sd, scale, n_res):
    net = ref.Net(num_channels=3, base_filter=256, num_residuals=n_res, upscale_factor=scale)
    assert list(net.state_dict().keys()) == list(sd.keys()), "EDSR key mismatch"
    net.load_state_dict(sd, strict=True)
    return net


def main():
    ref = ref_shim.load_reference("model.edsr")
    out = {}
    for name, scale, n_res, batch, lrs in EDSR_CASES:
        wseed, dseed = 50 + scale, 80 + scale
        sd = E.tie_upsampling(E.make_state(E.edsr_spec(scale, n_res), seed=wseed, init="fan"))
        net = build(ref, sd, scale, n_res)
        lr, hr = E.synthetic_batch(batch, scale, lrs * scale, seed=dseed)
        y = net(lr)
        loss = torch.nn.L1Loss()(y, hr)
        loss.backward()
        out[name] = {"cfg": dict(scale=scale, n_res=n_res, batch=batch, lr_size=lrs, wseed=wseed, dseed=dseed),
                     "out": y.detach().clone(), "loss": loss.item(),
                     "grads": {k: summarize(p.grad, 8) for k, p in net.named_parameters()}}
        print(name, tuple(y.shape), loss.item())
    c = TRAIN_CFG
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(c["scale"], c["n_res"]), seed=c["wseed"], init="fan"))
    net = build(ref, sd, c["scale"], c["n_res"])
    opt = torch.optim.Adam(net.parameters(), lr=c["lr"], betas=(0.9, 0.999))       # model/edsr.py:184
    crit = torch.nn.L1Loss()                                                          # :163-164
    steps = []
    for it in range(c["steps"]):
        lr, hr = E.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        opt.zero_grad()                                                               # :252
        gen_hr = net(lr)                                                              # :255
        loss = crit(gen_hr, hr)                                                       # :257-260
        loss.backward()                                                               # :264
        opt.step()                                                                    # :265
        steps.append({"loss_G": loss.item(), "params": {k: summarize(p, 8) for k, p in net.named_parameters()}})
        print("step", it, loss.item())
    out["train_steps"] = {"cfg": c, "steps": steps}
    path = os.path.join(GOLDEN_DIR, "edsr_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
