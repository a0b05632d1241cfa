"""CPU, build container only: pins oracle/sradsgan_oracle.py directly against the UNMODIFIED reference
classes imported from /root/reference (skipped on the GPU box, where the committed golden vectors of
tests/test_oracle_golden.py carry the pin)."""
import pytest
import torch

from oracle import ref_shim
from oracle import sradsgan_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


@pytest.mark.parametrize("scale", [2, 3, 4, 8, 9])
def test_state_dict_keys_and_shapes_match_reference(ref, scale):
    g = ref.GeneratorResNet(ref.ResGroup, n_residual_blocks=12, n_basic_blocks=3, upscale_factor=scale)
    spec = O.generator_spec(scale)
    sd = g.state_dict()
    assert list(sd.keys()) == list(spec.keys())
    assert all(tuple(sd[k].shape) == tuple(spec[k]) for k in spec)
    # weight tying of the upsampler stages (model/sradsgan.py:388-392)
    r, n = O.upsample_stages(scale)
    if n > 1:
        assert sd["GAB_UP.upsampling.0.weight"].data_ptr() == sd["GAB_UP.upsampling.3.weight"].data_ptr()
    d = ref.Discriminator().state_dict()
    dspec = O.discriminator_spec()
    assert list(d.keys()) == list(dspec.keys())
    assert all(tuple(d[k].shape) == tuple(dspec[k]) for k in dspec)


def test_reference_init_distribution(ref):
    """weights_init_normal (utils/utils.py:97-114) == make_state(init='ref') in distribution."""
    d = ref.Discriminator()
    d.apply(ref.srutils_mod.weights_init_normal)
    sd = d.state_dict()
    mine = O.make_state(O.discriminator_spec(), seed=3, init="ref")
    for k in ("model.5.weight", "model.6.weight", "model.6.bias", "model.5.bias"):
        assert abs(sd[k].float().mean().item() - mine[k].mean().item()) < 5e-3
        assert abs(sd[k].float().std().item() - mine[k].std().item()) < 5e-3


def test_forward_backward_matches_reference_modules(ref):
    scale, ng, nb = 4, 2, 2
    sd = O.tie_upsampling(O.make_state(O.generator_spec(scale, ng, nb), seed=77, init="fan"))
    G = ref.GeneratorResNet(ref.ResGroup, n_residual_blocks=ng, n_basic_blocks=nb, upscale_factor=scale)
    G.load_state_dict(sd, strict=True)
    lr, hr = O.synthetic_batch(2, scale, 40, seed=3)
    y_ref = G(lr)
    (y_ref - hr).abs().mean().backward()
    mine = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.tie_upsampling(mine)
    y = O.generator_forward(mine, lr, scale, ng, nb)
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    (y - hr).abs().mean().backward()
    for k, p in G.named_parameters():
        torch.testing.assert_close(mine[k].grad, p.grad, rtol=1e-4, atol=1e-6 + 1e-5 * p.grad.abs().max().item())
