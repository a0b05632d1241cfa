"""CPU: the SRAGAN sibling — oracle/sragan_oracle.py against the golden vectors recorded from the UNMODIFIED
reference `model.sragan` classes (oracle/make_golden_sragan.py) and, in the build container, against the imported reference itself;
then the product's host wiring (state_dict compatibility, BasicBlock = conv pair node + fused local-attention chain + activation, the
trainer's iteration) against the oracle with the C-ABI kernels replaced by oracle/ops_emu.py."""
import os

import numpy as np
import pytest
import torch

from oracle import ops_emu, ref_shim
from oracle import sradsgan_oracle as O
from oracle import sragan_oracle as A
from oracle.make_golden import summarize
from oracle.make_golden_sragan import SRAGAN_CASES
from sradsgan_b200 import _lib, ops
from sradsgan_b200.model.sragan import SRAGAN, BasicBlock, GeneratorResNet, ResidualBlock_Block_WithAttention
from test_srgan_cpu import srgan_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def agolden():
    return torch.load(os.path.join(ROOT, "tests", "golden", "sragan_golden.pt"), weights_only=False)


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


def sragan_args(**kw):
    base = dict(model_name="SRAGAN", gp=True, loss_Lp_norm="L1", weight_gan=1e-3, weight_content=1e-2)
    base.update(kw)
    return srgan_args(**base)


def _gen(scale, n_res, n_basic):
    return GeneratorResNet(ResidualBlock_Block_WithAttention, n_residual_blocks=n_res, n_basic_blocks=n_basic, upscale_factor=scale)


@pytest.mark.parametrize("case", SRAGAN_CASES, ids=lambda c: c[0])
def test_oracle_matches_reference_golden(agolden, case):
    name, scale, n_res, n_basic, batch, lrs = case
    g = agolden[name]
    sd = A.tie_upsampling(A.make_state(A.generator_spec(scale, n_res, n_basic), seed=g["cfg"]["wseed"], init="fan"))
    for p in A.unique_params(sd):
        p.requires_grad_(True)
    lr, hr = A.synthetic_batch(batch, scale, lrs * scale, seed=g["cfg"]["dseed"])
    y = A.generator_forward(sd, lr, scale, n_res, n_basic)
    torch.testing.assert_close(y.detach(), g["out"], rtol=1e-4, atol=1e-5 * g["out"].abs().max().item())
    loss = torch.nn.functional.l1_loss(y, hr)
    assert abs(loss.item() - g["loss"]) < 1e-5 * max(1.0, abs(g["loss"]))
    loss.backward()
    noise = A.noise_grad_keys(sd)
    for k, want in g["grads"].items():
        if k not in noise:
            assert abs(summarize(sd[k].grad, 8)["norm"] - want["norm"]) <= 2e-4 * max(1e-9, want["norm"]), k


def test_oracle_training_steps_match_reference_golden(agolden):
    c = agolden["train_steps"]["cfg"]
    G = A.tie_upsampling(A.make_state(A.generator_spec(c["scale"], c["n_res"], c["n_basic"]), seed=c["gseed"], init="fan"))
    D = O.make_state(O.discriminator_spec(), seed=c["dseed"], init="ref")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    st = A.TrainState(G, D, V, c["scale"], c["n_res"], c["n_basic"])
    noise = A.noise_grad_keys(G) | set(O.NOISE_GRAD_KEYS)
    for it, want in enumerate(agolden["train_steps"]["steps"]):
        lr, hr = A.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        np.random.seed(c["np_seed"] + it)
        out = A.train_step(st, lr, hr, torch.Tensor(np.random.random((c["batch"], 1, 1, 1))))
        for k in ("loss_G", "loss_D", "pixel", "content", "adv", "gp"):
            assert abs(out[k] - want[k]) <= 1e-4 * max(1.0, abs(want[k])), (it, k)
        for k, w in want["G_state"].items():
            if k not in noise:
                assert abs(summarize(G[k].float(), 8)["norm"] - w["norm"]) <= 1e-4 * max(1e-9, w["norm"]), (it, k)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
@pytest.mark.parametrize("scale", [2, 3, 4, 8, 9])
def test_oracle_and_product_keys_match_reference(scale):
    ref = ref_shim.load_reference("model.sragan")
    want = ref.GeneratorResNet(ref.ResidualBlock_Block_WithAttention, n_residual_blocks=2, n_basic_blocks=3, upscale_factor=scale).state_dict()
    spec = A.generator_spec(scale, 2, 3)
    assert list(want.keys()) == list(spec.keys())
    assert all(tuple(want[k].shape) == tuple(spec[k]) for k in spec)
    mine = _gen(scale, 2, 3).state_dict()
    assert list(mine.keys()) == list(want.keys())
    assert all(tuple(mine[k].shape) == tuple(want[k].shape) for k in want)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_oracle_forward_backward_matches_reference_module():
    ref = ref_shim.load_reference("model.sragan")
    scale, n_res, n_basic = 4, 2, 2
    sd = A.tie_upsampling(A.make_state(A.generator_spec(scale, n_res, n_basic), seed=5, init="fan"))
    net = ref.GeneratorResNet(ref.ResidualBlock_Block_WithAttention, n_residual_blocks=n_res, n_basic_blocks=n_basic, upscale_factor=scale)
    net.load_state_dict(sd, strict=True)
    net.train()
    lr, hr = A.synthetic_batch(2, scale, 40, seed=3)
    y_ref = net(lr)
    (y_ref - hr).abs().mean().backward()
    mine = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    A.tie_upsampling(mine)
    y = A.generator_forward(mine, lr, scale, n_res, n_basic)
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    (y - hr).abs().mean().backward()
    for k, p in net.named_parameters():
        torch.testing.assert_close(mine[k].grad, p.grad, rtol=1e-4, atol=1e-6 + 1e-5 * p.grad.abs().max().item())


def test_unbuilt_variants_are_refused():
    with pytest.raises(NotImplementedError):
        BasicBlock(64, 64)                                     # the reference's default norm_type='batch' is not part of SRAGAN
    with pytest.raises(NotImplementedError):
        BasicBlock(64, 64, norm_type=None, act_type='prelu')


@pytest.mark.parametrize("scale,n_basic", [(4, 2), (3, 1)])
def test_forward_backward_wiring(emu, scale, n_basic):
    n_res = 2
    sd = A.tie_upsampling(A.make_state(A.generator_spec(scale, n_res, n_basic), seed=7, init="fan"))
    net = _gen(scale, n_res, n_basic)
    net.load_state_dict(sd, strict=True)
    net.train()
    lr, hr = A.synthetic_batch(2, scale, 8 * scale, seed=9)
    y = net(lr)
    mine = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    A.tie_upsampling(mine)
    y_ref = A.generator_forward(mine, lr, scale, n_res, n_basic)
    assert rel(y, y_ref) < 1e-5
    (y.float() - hr).abs().mean().backward()
    (y_ref - hr).abs().mean().backward()
    for k, p in net.named_parameters():
        if mine[k].grad.abs().max() < 1e-7:
            assert p.grad is None or p.grad.abs().max() < 1e-6, k
            continue
        assert rel(p.grad, mine[k].grad) < 2e-3, k


def test_trainer_steps_match_reference_golden(emu, agolden):
    c = agolden["train_steps"]["cfg"]
    G = A.tie_upsampling(A.make_state(A.generator_spec(c["scale"], c["n_res"], c["n_basic"]), seed=c["gseed"], init="fan"))
    D = O.make_state(O.discriminator_spec(), seed=c["dseed"], init="ref")
    V = O.make_state(O.vgg_spec(), seed=c["vseed"], init="fan")
    net = SRAGAN(sragan_args(scale_factor=c["scale"], batch_size=c["batch"], vgg_state=V))
    net.n_residual_blocks, net.n_basic_blocks = c["n_res"], c["n_basic"]
    net.build(init=False)
    net.generator.load_state_dict(G, strict=True)
    net.discriminator.load_state_dict(D, strict=True)
    ops.bump_weight_generation()
    noise = A.noise_grad_keys(G) | set(O.NOISE_GRAD_KEYS)
    for it, want in enumerate(agolden["train_steps"]["steps"]):
        lr, hr = A.synthetic_batch(c["batch"], c["scale"], c["lr_size"] * c["scale"], seed=c["data_seed"] + it)
        np.random.seed(c["np_seed"] + it)
        net._alpha_override = torch.Tensor(np.random.random((c["batch"], 1, 1, 1)))
        out = net.train_step(lr, hr)
        for k in ("loss_G", "loss_D", "pixel", "content", "adv", "gp"):
            assert abs(out[k].item() - want[k]) <= 5e-4 * max(1.0, abs(want[k])), (it, k, out[k].item(), want[k])
        gsd = net.generator.state_dict()
        for k, w in want["G_state"].items():
            if k not in noise and "num_batches" not in k:
                assert abs(summarize(gsd[k].float(), 8)["norm"] - w["norm"]) <= 5e-4 * max(1e-9, w["norm"]), (it, k)


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_eval_mode_forward_matches_reference_module(emu):
    """validate() / inference run the generator in eval mode: BatchNorm with its running statistics == the reference module"""
    ref = ref_shim.load_reference("model.sragan")
    sd = A.make_state(A.generator_spec(4, 2, 2), seed=5, init="fan")
    g = torch.Generator().manual_seed(2)
    for k in sd:
        if "running_mean" in k:
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        if "running_var" in k:
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    A.tie_upsampling(sd)
    want = ref.GeneratorResNet(ref.ResidualBlock_Block_WithAttention, n_residual_blocks=2, n_basic_blocks=2, upscale_factor=4)
    want.load_state_dict(sd, strict=True)
    mine = _gen(4, 2, 2)
    mine.load_state_dict(sd, strict=True)
    x = torch.rand(2, 3, 10, 10, generator=g)
    with torch.no_grad():
        assert rel(mine.eval()(x), want.eval()(x)) < 1e-5
