"""CPU: the SRGAN sibling — oracle/srgan_oracle.py against the golden vectors recorded from the UNMODIFIED
reference `model.srgan` classes (oracle/make_golden_srgan.py) and, in the build container, against the imported reference itself;
then the product's host wiring (state_dict compatibility incl. the shared up-sampling conv / BatchNorm pair, the trainer's
G / D step bodies) against the oracle with the C-ABI kernels replaced by oracle/ops_emu.py."""
import os
import types

import pytest
import torch

from oracle import ops_emu, ref_shim
from oracle import sradsgan_oracle as O
from oracle import srgan_oracle as S
from oracle.make_golden import summarize
from oracle.make_golden_srgan import SRGAN_CASES
from sradsgan_b200 import _lib, ops
from sradsgan_b200.model.srgan import SRGAN, Discriminator, GeneratorResNet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sgolden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "srgan_golden.pt"), weights_only=False)


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


def srgan_args(**kw):
    base = dict(model_name="SRGAN", train_dataset=[], test_dataset=[], crop_size=32, test_crop_size=32, hr_height=32, hr_width=32,
                num_threads=0, num_channels=3, scale_factor=4, epoch=0, num_epochs=1, save_epochs=1, batch_size=2,
                test_batch_size=1, lr=2e-4, b1=0.9, b2=0.999, data_dir="", root_dir="", save_dir="/tmp/srgan_test", gpu_mode=True,
                n_cpu=0, sample_interval=1000, clip_value=0.01, lambda_gp=10, gp=False, penalty_type="LS",
                grad_penalty_Lp_norm="L2", relativeGan=False, loss_Lp_norm="L2", weight_gan=1e-3, weight_content=6e-3,
                max_train_samples=10, precision="fp32")
    base.update(kw)
    return types.SimpleNamespace(**base)


@pytest.mark.parametrize("case", SRGAN_CASES, ids=lambda c: c[0])
def test_oracle_matches_reference_golden(sgolden, case):
    name, scale, n_res, batch, lrs = case
    g = sgolden[name]
    sd = S.tie_upsampling(S.make_state(S.generator_spec(scale, n_res), seed=g["cfg"]["wseed"], init="fan"))
    for p in S.unique_params(sd):
        p.requires_grad_(True)
    lr, hr = S.synthetic_batch(batch, scale, lrs * scale, seed=g["cfg"]["dseed"])
    y = S.generator_forward(sd, lr, scale, n_res)
    torch.testing.assert_close(y.detach(), g["out"], rtol=1e-4, atol=1e-5 * g["out"].abs().max().item())
    loss = torch.nn.functional.mse_loss(y, hr)
    assert abs(loss.item() - g["loss"]) < 1e-5 * max(1.0, abs(g["loss"]))
    loss.backward()
    for k, want in g["grads"].items():
        assert abs(summarize(sd[k].grad, 8)["norm"] - want["norm"]) <= 2e-4 * max(1e-9, want["norm"]), k
    for k, want in g["buffers"].items():        # BatchNorm running statistics (the shared up-sampling BatchNorm is updated once per stage)
        assert abs(summarize(sd[k].float(), 8)["norm"] - want["norm"]) <= 1e-5 * max(1e-9, want["norm"]), k


def test_oracle_training_steps_match_reference_golden(sgolden):
    c = sgolden["train_steps"]["cfg"]
    G = S.tie_upsampling(S.make_state(S.generator_spec(c["scale"], c["n_res"]), seed=c["gseed"], init="fan"))
    D = S.make_state(S.discriminator_spec(), seed=c["dseed"], init="fan")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    st = S.TrainState(G, D, V, c["scale"], c["n_res"], lr=c["lr"])
    for it, want in enumerate(sgolden["train_steps"]["steps"]):
        lr, hr = S.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        out = S.train_step(st, lr, hr)
        for k in ("loss_G", "loss_D", "pixel", "content", "adv"):
            assert abs(out[k] - want[k]) <= 1e-4 * max(1.0, abs(want[k])), (it, k)
        for net, sd in (("G", G), ("D", D)):
            for k, w in want[net].items():
                assert abs(summarize(sd[k].float(), 8)["norm"] - w["norm"]) <= 1e-4 * max(1e-9, w["norm"]), (it, net, k)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
@pytest.mark.parametrize("scale", [2, 3, 4, 8, 9])
def test_oracle_and_product_keys_match_reference(scale):
    ref = ref_shim.load_reference("model.srgan")
    want = ref.GeneratorResNet(n_residual_blocks=2, upscale_factor=scale).state_dict()
    spec = S.generator_spec(scale, 2)
    assert list(want.keys()) == list(spec.keys())
    assert all(tuple(want[k].shape) == tuple(spec[k]) for k in spec)
    mine = GeneratorResNet(n_residual_blocks=2, upscale_factor=scale).state_dict()
    assert list(mine.keys()) == list(want.keys())
    assert all(tuple(mine[k].shape) == tuple(want[k].shape) for k in want)
    dwant = ref.Discriminator().state_dict()
    assert list(dwant.keys()) == list(S.discriminator_spec().keys()) == list(Discriminator().state_dict().keys())
    if scale in (4, 8, 9):      # the shared up-sampling conv / BatchNorm pair
        assert mine["upsampling.0.weight"].data_ptr() == mine["upsampling.4.weight"].data_ptr()
        assert mine["upsampling.1.running_mean"].data_ptr() == mine["upsampling.5.running_mean"].data_ptr()


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_oracle_forward_backward_matches_reference_module():
    ref = ref_shim.load_reference("model.srgan")
    scale, n_res = 4, 2
    sd = S.tie_upsampling(S.make_state(S.generator_spec(scale, n_res), seed=5, init="fan"))
    net = ref.GeneratorResNet(n_residual_blocks=n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.train()
    lr, hr = S.synthetic_batch(2, scale, 40, seed=3)
    y_ref = net(lr)
    ((y_ref - hr) ** 2).mean().backward()
    mine = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    S.tie_upsampling(mine)
    y = S.generator_forward(mine, lr, scale, n_res)
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    ((y - hr) ** 2).mean().backward()
    for k, p in net.named_parameters():
        torch.testing.assert_close(mine[k].grad, p.grad, rtol=1e-4, atol=1e-6 + 1e-5 * p.grad.abs().max().item())
    d = ref.Discriminator().train()
    dsd = S.make_state(S.discriminator_spec(), seed=6, init="fan")
    d.load_state_dict(dsd, strict=True)
    torch.testing.assert_close(S.discriminator_forward({k: v.clone() for k, v in dsd.items()}, hr), d(hr), rtol=1e-5, atol=1e-5)


def test_product_refuses_cpu():
    with pytest.raises(RuntimeError):
        GeneratorResNet(n_residual_blocks=1, upscale_factor=2)(torch.rand(1, 3, 8, 8))


@pytest.mark.parametrize("scale", [4, 3])
def test_forward_backward_wiring(emu, scale):
    n_res = 2
    sd = S.tie_upsampling(S.make_state(S.generator_spec(scale, n_res), seed=7, init="fan"))
    net = GeneratorResNet(n_residual_blocks=n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.train()
    lr, hr = S.synthetic_batch(2, scale, 8 * scale, seed=9)
    y = net(lr)
    mine = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    S.tie_upsampling(mine)
    y_ref = S.generator_forward(mine, lr, scale, n_res)
    assert rel(y, y_ref) < 1e-5
    ((y.float() - hr) ** 2).mean().backward()
    ((y_ref - hr) ** 2).mean().backward()
    for k, p in net.named_parameters():
        if mine[k].grad.abs().max() < 1e-7:      # a conv bias in front of a BatchNorm: its gradient is identically zero (rounding noise)
            assert p.grad.abs().max() < 1e-6, k
            continue
        assert rel(p.grad, mine[k].grad) < 2e-3, k
    gsd = net.state_dict()
    for k in sd:
        if "running" in k:
            torch.testing.assert_close(gsd[k], mine[k], rtol=1e-4, atol=1e-6)


def test_trainer_steps_match_reference_golden(emu, sgolden):
    c = sgolden["train_steps"]["cfg"]
    G = S.tie_upsampling(S.make_state(S.generator_spec(c["scale"], c["n_res"]), seed=c["gseed"], init="fan"))
    D = S.make_state(S.discriminator_spec(), seed=c["dseed"], init="fan")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    net = SRGAN(srgan_args(lr=c["lr"], scale_factor=c["scale"], batch_size=c["batch"], vgg_state=V))
    net.n_residual_blocks = c["n_res"]
    net.build(init=False)
    net.generator.load_state_dict(G, strict=True)
    net.discriminator.load_state_dict(D, strict=True)
    ops.bump_weight_generation()
    for it, want in enumerate(sgolden["train_steps"]["steps"]):
        lr, hr = S.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        out = net.train_step(lr, hr)
        for k in ("loss_G", "loss_D", "pixel", "content", "adv"):
            assert abs(out[k].item() - want[k]) <= 5e-4 * max(1.0, abs(want[k])), (it, k)
        for name, mod in (("G", net.generator), ("D", net.discriminator)):
            msd = mod.state_dict()
            noise = S.noise_grad_keys(msd)
            for k, w in want[name].items():
                if k in noise:
                    continue
                assert abs(summarize(msd[k].float(), 8)["norm"] - w["norm"]) <= 5e-4 * max(1e-9, w["norm"]), (it, name, k)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_eval_mode_forward_matches_reference_module(emu):
    """validate() / inference run the generator in eval mode: BatchNorm with its running statistics == the reference module"""
    ref = ref_shim.load_reference("model.srgan")
    sd = S.make_state(S.generator_spec(4, 2), seed=5, init="fan")
    g = torch.Generator().manual_seed(2)
    for k in sd:
        if "running_mean" in k:
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        if "running_var" in k:
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    S.tie_upsampling(sd)
    want = ref.GeneratorResNet(n_residual_blocks=2, upscale_factor=4)
    want.load_state_dict(sd, strict=True)
    mine = GeneratorResNet(n_residual_blocks=2, upscale_factor=4)
    mine.load_state_dict(sd, strict=True)
    x = torch.rand(2, 3, 10, 10, generator=g)
    with torch.no_grad():
        assert rel(mine.eval()(x), want.eval()(x)) < 1e-5
