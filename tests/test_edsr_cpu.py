"""CPU: the EDSR workload (BASELINE.json configs[4]) — oracle/edsr_oracle.py against the golden vectors
recorded from the UNMODIFIED reference `model.edsr.Net` (oracle/make_golden_edsr.py) and, in the build container, against
the imported reference itself; then the product's host wiring (module/state_dict compatibility, the ConvActConv node with
its residual epilogue, the EDSR trainer's step) against the oracle with the C-ABI kernels replaced by oracle/ops_emu.py."""
import os
import types

import pytest
import torch

from oracle import edsr_oracle as E
from oracle import ops_emu, ref_shim
from oracle.make_golden import summarize
from oracle.make_golden_edsr import EDSR_CASES
from sradsgan_b200 import _lib, ops
from sradsgan_b200.model.edsr import EDSR, ConvBlock, Net, ResnetBlock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def egolden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "edsr_golden.pt"), weights_only=False)


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


def edsr_args(**kw):
    base = dict(model_name="EDSR", train_dataset=[], test_dataset=[], crop_size=48, test_crop_size=48, hr_height=48, hr_width=48,
                num_threads=0, num_channels=3, scale_factor=4, epoch=0, num_epochs=1, save_epochs=1, batch_size=2,
                test_batch_size=1, lr=1e-4, b1=0.9, b2=0.999, data_dir="", root_dir="", save_dir="/tmp/edsr_test", gpu_mode=True,
                n_cpu=0, sample_interval=1000, clip_value=0.01, lambda_gp=10, gp=True, penalty_type="LS",
                grad_penalty_Lp_norm="L2", relativeGan=False, loss_Lp_norm="L1", weight_gan=1e-3, weight_content=1e-2,
                max_train_samples=10, precision="fp32")
    base.update(kw)
    return types.SimpleNamespace(**base)


@pytest.mark.parametrize("case", EDSR_CASES, ids=lambda c: c[0])
def test_oracle_matches_reference_golden(egolden, case):
    name, scale, n_res, batch, lrs = case
    g = egolden[name]
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(scale, n_res), seed=g["cfg"]["wseed"], init="fan"))
    for p in E.unique_params(sd):
        p.requires_grad_(True)
    lr, hr = E.synthetic_batch(batch, scale, lrs * scale, seed=g["cfg"]["dseed"])
    y = E.edsr_forward(sd, lr, scale, n_res)
    torch.testing.assert_close(y.detach(), g["out"], rtol=1e-4, atol=1e-5 * g["out"].abs().max().item())
    loss = torch.nn.functional.l1_loss(y, hr)
    assert abs(loss.item() - g["loss"]) < 1e-5 * max(1.0, abs(g["loss"]))
    loss.backward()
    for k, want in g["grads"].items():
        got = summarize(sd[k].grad, 8)
        assert abs(got["norm"] - want["norm"]) <= 1e-4 * max(1e-9, want["norm"]), k


def test_oracle_training_steps_match_reference_golden(egolden):
    c = egolden["train_steps"]["cfg"]
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(c["scale"], c["n_res"]), seed=c["wseed"], init="fan"))
    st = E.EdsrTrainState(sd, c["scale"], c["n_res"], lr=c["lr"])
    for it, want in enumerate(egolden["train_steps"]["steps"]):
        lr, hr = E.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        out = E.edsr_train_step(st, lr, hr)
        assert abs(out["loss_G"] - want["loss_G"]) <= 1e-4 * max(1.0, abs(want["loss_G"]))
        for k, w in want["params"].items():
            assert abs(summarize(sd[k], 8)["norm"] - w["norm"]) <= 1e-5 * max(1e-9, w["norm"]), (it, k)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
@pytest.mark.parametrize("scale", [2, 3, 4, 8, 9])
def test_oracle_and_product_keys_match_reference(scale):
    ref = ref_shim.load_reference("model.edsr")
    want = ref.Net(3, 256, 2, upscale_factor=scale).state_dict()
    spec = E.edsr_spec(scale, 2)
    assert list(want.keys()) == list(spec.keys())
    assert all(tuple(want[k].shape) == tuple(spec[k]) for k in spec)
    mine = Net(3, 256, 2, upscale_factor=scale).state_dict()
    assert list(mine.keys()) == list(want.keys())
    assert all(tuple(mine[k].shape) == tuple(want[k].shape) for k in want)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_oracle_forward_backward_matches_reference_module():
    ref = ref_shim.load_reference("model.edsr")
    scale, n_res = 4, 2
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(scale, n_res), seed=5, init="fan"))
    net = ref.Net(3, 256, n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    lr, hr = E.synthetic_batch(2, scale, 40, seed=3)
    y_ref = net(lr)
    (y_ref - hr).abs().mean().backward()
    mine = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    E.tie_upsampling(mine)
    y = E.edsr_forward(mine, lr, scale, n_res)
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    (y - hr).abs().mean().backward()
    for k, p in net.named_parameters():
        torch.testing.assert_close(mine[k].grad, p.grad, rtol=1e-4, atol=1e-6 + 1e-5 * p.grad.abs().max().item())


def test_state_dict_keys_and_tying():
    for scale in (2, 3, 4, 8, 9):
        net = Net(3, 256, 2, upscale_factor=scale)
        spec = E.edsr_spec(scale, 2)
        sd = net.state_dict()
        assert list(sd.keys()) == list(spec.keys())
        assert all(tuple(sd[k].shape) == tuple(v) for k, v in spec.items())
        if "upsampling.3.weight" in sd:
            assert sd["upsampling.0.weight"].data_ptr() == sd["upsampling.3.weight"].data_ptr()
    with pytest.raises(NotImplementedError):
        ResnetBlock(256)                 # the reference's default norm='batch' is not part of EDSR
    with pytest.raises(NotImplementedError):
        ConvBlock(3, 8, 3, 1, 1, norm='batch')


def test_product_refuses_cpu():
    with pytest.raises(RuntimeError):
        Net(3, 256, 1, upscale_factor=2)(torch.rand(1, 3, 8, 8))


@pytest.mark.parametrize("scale", [4, 3])
def test_forward_backward_wiring(emu, scale):
    n_res = 2
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(scale, n_res), seed=7, init="fan"))
    net = Net(3, 256, n_res, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    lr, hr = E.synthetic_batch(2, scale, 8 * scale, seed=9)
    y = net(lr)
    mine = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    E.tie_upsampling(mine)
    y_ref = E.edsr_forward(mine, lr, scale, n_res)
    assert rel(y, y_ref) < 1e-5
    (y.float() - hr).abs().mean().backward()
    (y_ref - hr).abs().mean().backward()
    for k, p in net.named_parameters():
        assert rel(p.grad, mine[k].grad) < 2e-3, k      # fp32 summation order through the 256-channel trunk


def test_trainer_steps_match_reference_golden(emu, egolden):
    c = egolden["train_steps"]["cfg"]
    sd = E.tie_upsampling(E.make_state(E.edsr_spec(c["scale"], c["n_res"]), seed=c["wseed"], init="fan"))
    net = EDSR(edsr_args(lr=c["lr"], scale_factor=c["scale"], batch_size=c["batch"]))
    net.num_residuals = c["n_res"]
    net.build(init=False)
    net.generator.load_state_dict(sd, strict=True)
    ops.bump_weight_generation()
    for it, want in enumerate(egolden["train_steps"]["steps"]):
        lr, hr = E.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        out = net.train_step(lr, hr)
        assert abs(out["loss_G"].item() - want["loss_G"]) <= 2e-4 * max(1.0, abs(want["loss_G"]))
        gsd = net.generator.state_dict()
        for k, w in want["params"].items():
            assert abs(summarize(gsd[k], 8)["norm"] - w["norm"]) <= 2e-4 * max(1e-9, w["norm"]), (it, k)
