"""CPU: host-side logic of the product — module/state_dict compatibility with the reference, the autograd
wiring of sradsgan_b200/ops.py (incl. the double-backward closure of the conv trio), the fused-Adam flat
buffers and the trainer's step structure — checked against the oracle with the C-ABI kernels replaced by
their documented-semantics emulation (oracle/ops_emu.py).  No CUDA compute happens here; the kernels
themselves are checked by the `-m gpu` tests."""
import types

import numpy as np
import pytest
import torch

from oracle import ops_emu
from oracle import sradsgan_oracle as O
from sradsgan_b200 import _lib, ops
from sradsgan_b200.model.sradsgan import (Discriminator, FeatureExtractor, GeneratorResNet, ResGroup, SRADSGAN)


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


@pytest.mark.parametrize("scale", [2, 3, 4, 8, 9])
def test_state_dict_keys_shapes_and_tying(scale):
    g = GeneratorResNet(ResGroup, upscale_factor=scale)
    spec = O.generator_spec(scale)
    sd = g.state_dict()
    assert list(sd.keys()) == list(spec.keys())
    assert all(tuple(sd[k].shape) == tuple(v) for k, v in spec.items())
    if "GAB_UP.upsampling.3.weight" in sd:
        assert sd["GAB_UP.upsampling.0.weight"].data_ptr() == sd["GAB_UP.upsampling.3.weight"].data_ptr()
    d = Discriminator().state_dict()
    dspec = O.discriminator_spec()
    assert list(d.keys()) == list(dspec.keys())
    assert all(tuple(d[k].shape) == tuple(v) for k, v in dspec.items())
    assert list(FeatureExtractor().state_dict().keys()) == list(O.vgg_spec().keys())


def test_product_refuses_cpu_without_extension_path():
    g = GeneratorResNet(ResGroup, n_residual_blocks=1, n_basic_blocks=1)
    with pytest.raises(RuntimeError):
        g(torch.rand(1, 3, 8, 8))


@pytest.mark.parametrize("scale", [4, 3])
def test_generator_forward_backward_wiring(emu, scale):
    ng, nb = 2, 1
    sd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=5, init="fan"))
    G = GeneratorResNet(ResGroup, n_residual_blocks=ng, n_basic_blocks=nb, upscale_factor=scale)
    G.load_state_dict(sd, strict=True)
    lr, hr = O.synthetic_batch(2, scale, 8 * scale, seed=9)
    y = G(lr)
    mine = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.tie_upsampling(mine)
    y_ref = O.generator_forward(mine, lr, scale, ng, nb)
    assert rel(y, y_ref) < 1e-5
    (y.float() - hr).abs().mean().backward()
    (y_ref - hr).abs().mean().backward()
    for k, p in G.named_parameters():
        if k in O.NOISE_GRAD_KEYS:
            continue
        assert rel(p.grad, mine[k].grad) < 2e-4, k


def test_discriminator_double_backward_wiring(emu):
    sd = O.make_state(O.discriminator_spec(), seed=3, init="fan")
    D = Discriminator()
    D.load_state_dict(sd, strict=True)
    ref = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    x = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(1))

    def gp_of(fwd):
        xi = x.clone().requires_grad_(True)
        d = fwd(xi)
        g = torch.autograd.grad(d, xi, torch.ones_like(d), create_graph=True, retain_graph=True)[0]
        return ((g.float().norm(2, 1) - 1) ** 2).mean()

    gp = gp_of(D)
    gp_ref = gp_of(lambda t: O.discriminator_forward(ref, t))
    assert abs(gp.item() - gp_ref.item()) < 1e-5 * max(1.0, abs(gp_ref.item()))
    gp.backward()
    gp_ref.backward()
    for k, p in D.named_parameters():
        if k in O.NOISE_GRAD_KEYS:
            continue
        if ref[k].grad is None:      # e.g. the last bias: the input-gradient does not depend on it
            assert p.grad is None or p.grad.abs().max() == 0, k
            continue
        if p.grad is None:           # BatchNorm shifts: the penalty's gradient does not depend on them (exact zeros in torch)
            assert ref[k].grad.abs().max() == 0, k
            continue
        assert rel(p.grad, ref[k].grad) < 5e-4, k
    for k in sd:
        if "running" in k:
            torch.testing.assert_close(D.state_dict()[k], ref[k], rtol=1e-4, atol=1e-6)


def _args(**kw):
    base = dict(model_name="SRADSGAN", train_dataset=[], test_dataset=[], crop_size=32, test_crop_size=32, hr_height=32,
                hr_width=32, num_threads=0, num_channels=3, scale_factor=4, epoch=0, num_epochs=1, save_epochs=1,
                batch_size=2, test_batch_size=1, lr=2e-4, b1=0.9, b2=0.999, data_dir="", root_dir="", save_dir="/tmp/sr_t",
                gpu_mode=True, n_cpu=0, sample_interval=1000, clip_value=0.01, lambda_gp=10, gp=True, penalty_type="LS",
                grad_penalty_Lp_norm="L2", relativeGan=False, loss_Lp_norm="L1", weight_gan=1e-3, weight_content=1e-2,
                max_train_samples=10, precision="fp32")
    base.update(kw)
    return types.SimpleNamespace(**base)


@pytest.mark.parametrize("fuse", [True, False])
def test_train_step_matches_oracle_and_golden(emu, golden, fuse):
    """two full G+D iterations through the product trainer == oracle == reference golden"""
    gcfg = golden["train_steps"]["cfg"]
    ng, nb, scale = gcfg["n_groups"], gcfg["n_blocks"], gcfg["scale"]
    Gsd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=gcfg["gseed"], init="fan"))
    Dsd = O.make_state(O.discriminator_spec(), seed=gcfg["dseed"], init="ref")
    Vsd = O.make_state(O.vgg_spec(), seed=gcfg["vseed"], init="fan")
    net = SRADSGAN(_args(vgg_state=Vsd))
    net.new_generator = lambda: GeneratorResNet(ResGroup, n_residual_blocks=ng, n_basic_blocks=nb, upscale_factor=scale)
    net.build(init=False)
    net.generator.load_state_dict(Gsd, strict=True)
    net.discriminator.load_state_dict(Dsd, strict=True)
    ops.bump_weight_generation()
    for it, want in enumerate(golden["train_steps"]["steps"]):
        lr, hr = O.synthetic_batch(gcfg["batch"], scale, gcfg["lr_size"] * scale, seed=gcfg["data_seed"] + it)
        np.random.seed(gcfg["np_seed"] + it)
        net._alpha_override = torch.Tensor(np.random.random((gcfg["batch"], 1, 1, 1)))
        out = net.train_step(lr, hr, fuse_gp_backward=fuse)
        for k in ("loss_G", "loss_D", "pixel", "content", "adv", "gp"):
            assert abs(out[k].item() - want[k]) <= 3e-4 * max(1.0, abs(want[k])), (it, k, out[k].item(), want[k])
        from oracle.make_golden import summarize
        gsd = net.generator.state_dict()
        for k, w in want["G_params"].items():
            if k in O.NOISE_GRAD_KEYS:
                continue
            assert abs(summarize(gsd[k], 8)["norm"] - w["norm"]) <= 2e-4 * max(1e-6, w["norm"]), (it, k)
        dsd = net.discriminator.state_dict()
        for k, w in want["D_state"].items():
            if k in O.NOISE_GRAD_KEYS:
                continue
            assert abs(summarize(dsd[k].float(), 8)["norm"] - w["norm"]) <= 2e-3 * max(1e-6, w["norm"]), (it, k)


def test_device_side_metrics_match_the_reference_definitions():
    """utils.psnr_batch / ergas_batch / quantize_u8 == the per-image host definitions of the reference
    (utils/utils.py:700-709 PSNR, :954-962 compare_ergas2 on the uint8 images, :169-175 save_img1 truncation)."""
    from sradsgan_b200 import utils as U
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(3, 3, 20, 24, generator=g)
    pred = gt + 0.05 * torch.randn(3, 3, 20, 24, generator=g)
    pred[2] = gt[2]                                            # zero-MSE image -> 100 dB
    got = U.psnr_batch(pred, gt)
    for j in range(3):
        assert abs(got[j].item() - O.psnr(pred[j], gt[j])) < 1e-9
        assert abs(got[j].item() - U.psnr(pred[j], gt[j])) < 1e-9
    e = U.ergas_batch(pred, gt, scale=4)
    for j in range(3):
        a = O.quantize_u8(gt[j]).astype(np.float64); b = O.quantize_u8(pred[j]).astype(np.float64)
        want = 100.0 * np.sqrt(np.mean((a - b) ** 2) / np.mean(a) ** 2 / 3) / 4
        assert abs(e[j].item() - want) < 1e-9 * max(1.0, want)
        assert np.array_equal(U.quantize_u8(pred[j]), O.quantize_u8(pred[j]))


def test_band_path_wiring_equals_tile_path_wiring():
    """the chain's band-path extras — the dense-sampling accumulator `out_all += y` done inside the chain (reference :459) and
    the pooling partials handed from producer to consumer — change the autograd graph (extra Function inputs / outputs, the
    accumulator's gradient passing through), not the arithmetic: with the emulated backend both wirings must give identical
    outputs and gradients."""
    prev_dtype = ops.config.compute_dtype
    res = []
    try:
        for band in (False, True):
            prev = _lib.set_backend(ops_emu.EmuBackend(band=band))
            ops.set_precision("bf16")
            try:
                scale, ng, nb = 4, 3, 2
                sd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=5, init="fan"))
                G = GeneratorResNet(ResGroup, n_residual_blocks=ng, n_basic_blocks=nb, upscale_factor=scale)
                G.load_state_dict(sd, strict=True)
                lr, hr = O.synthetic_batch(2, scale, 8 * scale, seed=9)
                y = G(lr)
                (0.5 * (y.float() - hr) ** 2).mean().backward()
                res.append((y.detach().clone(), {k: p.grad.clone() for k, p in G.named_parameters()}))
            finally:
                _lib.set_backend(prev)
    finally:
        ops.config.compute_dtype = prev_dtype
    (y0, g0), (y1, g1) = res
    assert torch.equal(y0, y1)
    for k in g0:
        assert torch.allclose(g0[k], g1[k], rtol=1e-5, atol=1e-8), k
