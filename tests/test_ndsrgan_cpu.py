"""CPU: the NDSRGAN sibling — oracle/ndsrgan_oracle.py against the golden vectors recorded from the UNMODIFIED
reference `model.ndsrgan` classes (oracle/make_golden_ndsrgan.py) and, in the build container, against the imported reference itself;
then the product's host wiring (state_dict compatibility incl. the shared up-sampling conv, the densely connected trunk, the trainer's
Smooth-L1 iteration) against the oracle with the C-ABI kernels replaced by oracle/ops_emu.py."""
import os

import pytest
import torch

from oracle import ndsrgan_oracle as N
from oracle import ops_emu, ref_shim
from oracle import sradsgan_oracle as O
from oracle.make_golden import summarize
from oracle.make_golden_ndsrgan import NDSRGAN_CASES
from sradsgan_b200 import _lib, ops
from sradsgan_b200.model.ndsrgan import NDSRGAN, Discriminator, GeneratorResNet
from test_srgan_cpu import srgan_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ngolden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "ndsrgan_golden.pt"), weights_only=False)


@pytest.fixture()
def emu():
    prev = _lib.set_backend(ops_emu.EmuBackend())
    prev_dtype = ops.config.compute_dtype
    ops.set_precision("fp32")
    yield
    ops.config.compute_dtype = prev_dtype
    _lib.set_backend(prev)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("case", NDSRGAN_CASES, ids=lambda c: c[0])
def test_oracle_matches_reference_golden(ngolden, case):
    name, scale, batch, lrs = case
    g = ngolden[name]
    sd = N.make_gen_state(scale, 23, g["cfg"]["wseed"])
    for p in N.unique_params(sd):
        p.requires_grad_(True)
    lr, hr = N.synthetic_batch(batch, scale, lrs * scale, seed=g["cfg"]["dseed"])
    y = N.generator_forward(sd, lr, scale)
    torch.testing.assert_close(y.detach(), g["out"], rtol=1e-4, atol=1e-5 * g["out"].abs().max().item())
    loss = torch.nn.functional.smooth_l1_loss(y, hr)
    assert abs(loss.item() - g["loss"]) < 1e-5 * max(1.0, abs(g["loss"]))
    if g["grads"]:
        loss.backward()
        for k, want in g["grads"].items():
            assert abs(summarize(sd[k].grad, 8)["norm"] - want["norm"]) <= 2e-4 * max(1e-9, want["norm"]), k


def test_oracle_training_steps_match_reference_golden(ngolden):
    c = ngolden["train_steps"]["cfg"]
    G = N.make_gen_state(c["scale"], 23, c["gseed"])
    D = N.make_state(N.discriminator_spec(), seed=c["dseed"], init="fan")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    st = N.TrainState(G, D, V, c["scale"], 23, lr=c["lr"])
    for it, want in enumerate(ngolden["train_steps"]["steps"]):
        lr, hr = N.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        out = N.train_step(st, lr, hr)
        for k in ("loss_G", "loss_D", "pixel", "content", "adv"):
            assert abs(out[k] - want[k]) <= 1e-4 * max(1.0, abs(want[k])), (it, k)
        for k, w in want["G"].items():
            assert abs(summarize(G[k].float(), 8)["norm"] - w["norm"]) <= 1e-4 * max(1e-9, w["norm"]), (it, k)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
@pytest.mark.parametrize("scale", [2, 3, 4, 8, 9])
def test_oracle_and_product_keys_match_reference(scale):
    ref = ref_shim.load_reference("model.ndsrgan")
    want = ref.GeneratorResNet(upscale_factor=scale).state_dict()
    spec = N.generator_spec(scale)
    assert list(want.keys()) == list(spec.keys())
    assert all(tuple(want[k].shape) == tuple(spec[k]) for k in spec)
    mine = GeneratorResNet(upscale_factor=scale).state_dict()
    assert list(mine.keys()) == list(want.keys())
    assert all(tuple(mine[k].shape) == tuple(want[k].shape) for k in want)
    assert list(ref.Discriminator().state_dict().keys()) == list(N.discriminator_spec().keys()) == list(Discriminator().state_dict().keys())
    if scale in (4, 8, 9):
        assert mine["upsampling.1.weight"].data_ptr() == mine["upsampling.4.weight"].data_ptr()


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_oracle_critic_matches_reference_module():
    ref = ref_shim.load_reference("model.ndsrgan")
    d = ref.Discriminator().train()
    dsd = N.make_state(N.discriminator_spec(), seed=6, init="fan")
    d.load_state_dict(dsd, strict=True)
    x = torch.rand(2, 3, 40, 40, generator=torch.Generator().manual_seed(1))
    torch.testing.assert_close(N.discriminator_forward({k: v.clone() for k, v in dsd.items()}, x), d(x), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("scale", [4, 3])
def test_forward_backward_wiring(emu, scale):
    n_blocks = 2
    sd = N.make_gen_state(scale, n_blocks, 7)
    net = GeneratorResNet(upscale_factor=scale, n_blocks=n_blocks)
    net.load_state_dict(sd, strict=True)
    net.train()
    lr, hr = N.synthetic_batch(2, scale, 6 * scale, seed=9)
    y = net(lr)
    mine = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    N.tie_upsampling(mine)
    y_ref = N.generator_forward(mine, lr, scale, n_blocks)
    assert rel(y, y_ref) < 1e-5
    ((y.float() - hr) ** 2).mean().backward()
    ((y_ref - hr) ** 2).mean().backward()
    for k, p in net.named_parameters():
        assert rel(p.grad, mine[k].grad) < 2e-3, k


def test_critic_wiring(emu):
    dsd = N.make_state(N.discriminator_spec(), seed=6, init="fan")
    D = Discriminator()
    D.load_state_dict(dsd, strict=True)
    D.train()
    x = torch.rand(2, 3, 40, 40, generator=torch.Generator().manual_seed(1))
    ref = {k: v.clone() for k, v in dsd.items()}
    assert rel(D(x), N.discriminator_forward(ref, x)) < 1e-5
    for k in dsd:
        if "running" in k:
            torch.testing.assert_close(D.state_dict()[k], ref[k], rtol=1e-4, atol=1e-6)


def test_trainer_steps_match_reference_golden(emu, ngolden):
    c = ngolden["train_steps"]["cfg"]
    G = N.make_gen_state(c["scale"], 23, c["gseed"])
    D = N.make_state(N.discriminator_spec(), seed=c["dseed"], init="fan")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    net = NDSRGAN(srgan_args(model_name="NDSRGAN", lr=c["lr"], scale_factor=c["scale"], batch_size=c["batch"], vgg_state=V))
    net.build(init=False)
    net.generator.load_state_dict(G, strict=True)
    net.discriminator.load_state_dict(D, strict=True)
    ops.bump_weight_generation()
    noise = N.noise_grad_keys(net.discriminator.state_dict())
    for it, want in enumerate(ngolden["train_steps"]["steps"]):
        lr, hr = N.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        out = net.train_step(lr, hr)
        for k in ("loss_G", "loss_D", "pixel", "content", "adv"):
            assert abs(out[k].item() - want[k]) <= 5e-4 * max(1.0, abs(want[k])), (it, k, out[k].item(), want[k])
        gsd = net.generator.state_dict()
        for k, w in want["G"].items():
            assert abs(summarize(gsd[k].float(), 8)["norm"] - w["norm"]) <= 5e-4 * max(1e-9, w["norm"]), (it, k)
        dsd = net.discriminator.state_dict()
        for k, w in want["D"].items():
            if k not in noise and "num_batches" not in k:
                assert abs(summarize(dsd[k].float(), 8)["norm"] - w["norm"]) <= 2e-3 * max(1e-9, w["norm"]), (it, k)
