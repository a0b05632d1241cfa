"""CPU: the restated oracle (oracle/sradsgan_oracle.py) reproduces the golden vectors that
oracle/make_golden.py recorded from the UNMODIFIED reference classes."""
import numpy as np
import pytest
import torch

from oracle import sradsgan_oracle as O
from oracle.make_golden import GEN_CASES, summarize


def _close_summary(got, want, rtol=2e-5):
    s = summarize(got, want["samples"].numel())
    assert s["shape"] == tuple(want["shape"])
    assert abs(s["norm"] - want["norm"]) <= rtol * max(1e-6, abs(want["norm"]))
    torch.testing.assert_close(s["samples"], want["samples"], rtol=1e-4, atol=rtol * max(1e-6, want["norm"] / max(1, got.numel()) ** 0.5) * 10)


@pytest.mark.parametrize("case", [c for c in GEN_CASES if not c[0].startswith("g_x4_full")], ids=lambda c: c[0])
def test_generator_forward_matches_reference_golden(golden, case):
    name, scale, ng, nb, batch, lrs, init = case
    g = golden[name]
    sd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=g["cfg"]["wseed"], init=init))
    lr, hr = O.synthetic_batch(batch, scale, lrs * scale, seed=g["cfg"]["dseed"])
    taps = {}
    with torch.no_grad():
        y = O.generator_forward(sd, lr, scale, ng, nb, taps)
    torch.testing.assert_close(y, g["out"], rtol=1e-4, atol=1e-5 * g["out"].abs().max().item())
    assert abs(O.psnr(y, hr) - g["psnr_vs_hr"]) < 1e-4
    checked = 0
    for k, v in taps.items():
        if k in g["taps"]:
            _close_summary(v, g["taps"][k])
            checked += 1
    assert checked >= 6


@pytest.mark.parametrize("cname", ["g_x4_full", "g_x4_full_refinit"])
def test_generator_full_architecture_golden(golden, cname):
    name, scale, ng, nb, batch, lrs, init = [c for c in GEN_CASES if c[0] == cname][0]
    g = golden[name]
    sd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=g["cfg"]["wseed"], init=init))
    n_unique = sum(p.numel() for p in O.unique_params(sd))
    assert n_unique == 11069493
    lr, _ = O.synthetic_batch(batch, scale, lrs * scale, seed=g["cfg"]["dseed"])
    with torch.no_grad():
        y = O.generator_forward(sd, lr, scale, ng, nb)
    torch.testing.assert_close(y, g["out"], rtol=1e-4, atol=2e-5 * g["out"].abs().max().item())


def test_discriminator_forward_and_bn_stats_golden(golden):
    g = golden["d_fwd"]
    sd = O.make_state(O.discriminator_spec(), seed=11, init="fan")
    assert sum(p.numel() for p in O.unique_params(sd)) == 4701987
    x = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(5))
    taps = {}
    with torch.no_grad():
        y = O.discriminator_forward(sd, x, update_stats=True, taps=taps)
    torch.testing.assert_close(y, g["out"], rtol=1e-4, atol=1e-6)
    for k, v in g["bn"].items():
        torch.testing.assert_close(sd[k].to(v.dtype), v, rtol=1e-5, atol=1e-7)
    for k, v in g["taps"].items():
        if k in taps:
            _close_summary(taps[k], v)


def test_vgg_golden(golden):
    sd = O.make_state(O.vgg_spec(), seed=12, init="fan")
    x = torch.rand(1, 3, 16, 16, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        y = O.vgg_features(sd, x)
    torch.testing.assert_close(y, golden["vgg"]["out"], rtol=1e-4, atol=1e-6)


def test_two_training_iterations_golden(golden):
    """losses, gradients, post-step parameters and D's BN buffers after 2 iterations (incl. the
    double-counted gradient penalty and the weight clamp)."""
    g = golden["train_steps"]
    c = g["cfg"]
    G = O.tie_upsampling(O.make_state(O.generator_spec(c["scale"], c["n_groups"], c["n_blocks"]), seed=c["gseed"], init="fan"))
    D = O.make_state(O.discriminator_spec(), seed=c["dseed"], init="ref")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    st = O.TrainState(G, D, V, c["scale"], c["n_groups"], c["n_blocks"])
    for it, want in enumerate(g["steps"]):
        lr, hr = O.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        np.random.seed(c["np_seed"] + it)
        alpha = torch.Tensor(np.random.random((c["batch"], 1, 1, 1)))
        rec = O.train_step(st, lr, hr, alpha)
        for k in ("loss_G", "loss_D", "pixel", "content", "adv", "gp"):
            assert abs(rec[k] - want[k]) <= 2e-4 * max(1.0, abs(want[k])), (it, k, rec[k], want[k])
        for k, w in want["G_params"].items():
            if k not in O.NOISE_GRAD_KEYS:
                _close_summary(G[k], w, rtol=1e-4)
        for k, w in want["D_state"].items():
            if k not in O.NOISE_GRAD_KEYS:
                _close_summary(D[k].float(), w, rtol=1e-4)
        for k, w in want["D_grads"].items():
            if k in O.NOISE_GRAD_KEYS:
                assert D[k].grad.abs().max().item() < 1e-4      # exact gradient is zero (fp noise only)
                continue
            got = summarize(D[k].grad, 8)
            assert abs(got["norm"] - w["norm"]) <= 2e-3 * max(1e-7, w["norm"]), (it, k)


def test_full_size_iteration_golden():
    """the oracle at the BENCHMARKED configuration (full 12x3 G + D + VGG19[:12], batch 16, LR 54^2 / HR 216^2) == one
    iteration of the unmodified reference recorded by `python -m oracle.make_golden --fullsize` (≈40 s of CPU work): losses,
    generator-output PSNR, gradients and post-step parameters.  The GPU suite holds `graphed_step` to the same numbers."""
    import os
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sradsgan_fullsize_golden.pt"), weights_only=False)
    c, want = g["cfg"], g["step"]
    torch.set_num_threads(os.cpu_count() or 1)
    G = O.tie_upsampling(O.make_state(O.generator_spec(c["scale"]), seed=c["gseed"], init="ref", gamma=c["gamma"]))
    D = O.make_state(O.discriminator_spec(), seed=c["dseed"], init="ref")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    lr, hr = O.synthetic_batch(c["batch"], c["scale"], c["hr"], seed=c["data_seed"])
    st = O.TrainState(G, D, V, c["scale"])
    np.random.seed(c["np_seed"])
    alpha = torch.from_numpy(np.random.random((c["batch"], 1, 1, 1))).float()
    rec = O.train_step(st, lr, hr, alpha)
    for k in ("loss_G", "loss_D", "pixel", "content", "adv", "gp"):
        assert abs(rec[k] - want[k]) <= 1e-5 * max(1.0, abs(want[k])), (k, rec[k], want[k])
    _close_summary(rec["gen_hr"], want["gen_hr"], rtol=1e-5)
    assert abs(O.psnr(rec["gen_hr"], hr) - want["psnr_vs_hr"]) < 1e-4
    for k, w in want["G_grads"].items():
        if k not in O.NOISE_GRAD_KEYS and k in G:
            got = summarize(G[k].grad, 8)
            assert abs(got["norm"] - w["norm"]) <= 1e-3 * max(1e-12, w["norm"]), k
    for k, w in want["D_grads"].items():
        if k not in O.NOISE_GRAD_KEYS:
            got = summarize(D[k].grad, 8)
            assert abs(got["norm"] - w["norm"]) <= 2e-3 * max(1e-12, w["norm"]), k
    for k, w in want["D_state"].items():
        if "running" in k:
            _close_summary(D[k].float(), w, rtol=1e-5)
