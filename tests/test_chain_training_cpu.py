"""CPU: chain training (BASELINE.json configs[2]; reference model/sradsgan.py:716-721) — x2 -> x3 -> x4 stages in one
process, each a complete `train()` whose generator / critic are warm-started from the previous stage.  The C-ABI kernels
are replaced by oracle/ops_emu.py (host logic only); the kernels are covered by the `-m gpu` tests at every scale."""
import types

import pytest
import torch

from oracle import ops_emu
from oracle import sradsgan_oracle as O
from sradsgan_b200 import _lib, ops
from sradsgan_b200.model.sradsgan import GeneratorResNet, ResGroup, SRADSGAN


@pytest.fixture()
def emu():
    prev = _lib.set_backend(ops_emu.EmuBackend())
    prev_dtype = ops.config.compute_dtype
    ops.set_precision("fp32")
    yield
    ops.config.compute_dtype = prev_dtype
    _lib.set_backend(prev)


def _args(tmp, **kw):
    base = dict(model_name="SRADSGAN", train_dataset=[], test_dataset=[], crop_size=36, test_crop_size=36, hr_height=36, hr_width=36,
                num_threads=0, num_channels=3, scale_factor=2, epoch=0, num_epochs=1, save_epochs=1, batch_size=1,
                test_batch_size=1, lr=2e-4, b1=0.9, b2=0.999, data_dir="", root_dir="", save_dir=str(tmp), gpu_mode=True,
                n_cpu=0, sample_interval=1000, clip_value=0.01, lambda_gp=10, gp=True, penalty_type="LS",
                grad_penalty_Lp_norm="L2", relativeGan=False, loss_Lp_norm="L1", weight_gan=1e-3, weight_content=1e-2,
                max_train_samples=10, precision="fp32", synthetic_steps=2, vgg_state=O.make_state(O.vgg_spec(), seed=2, init="fan"))
    base.update(kw)
    return types.SimpleNamespace(**base)


class Recording(SRADSGAN):
    def new_generator(self):
        return GeneratorResNet(ResGroup, n_residual_blocks=1, n_basic_blocks=1, upscale_factor=self.scale_factor)

    def load_pretrained(self, G_path=None, D_path=None):
        super().load_pretrained(G_path, D_path)
        self.started_from = ({k: v.detach().clone() for k, v in self.generator.state_dict().items()},
                             {k: v.detach().clone() for k, v in self.discriminator.state_dict().items()})


def test_warm_start_skips_only_mismatched_shapes():
    g2 = GeneratorResNet(ResGroup, n_residual_blocks=1, n_basic_blocks=1, upscale_factor=2)
    g3 = GeneratorResNet(ResGroup, n_residual_blocks=1, n_basic_blocks=1, upscale_factor=3)
    g4 = GeneratorResNet(ResGroup, n_residual_blocks=1, n_basic_blocks=1, upscale_factor=4)
    loaded, skipped = SRADSGAN.warm_start(g3, g2.state_dict())
    assert skipped == ["GAB_UP.upsampling.0.weight", "GAB_UP.upsampling.0.bias"]
    assert all(torch.equal(g3.state_dict()[k], g2.state_dict()[k]) for k in loaded)
    # x2 -> x4: same family, every entry of x2 carries over; x4's second (aliased) stage shares the loaded tensor
    loaded, skipped = SRADSGAN.warm_start(g4, g2.state_dict())
    assert skipped == []
    sd4 = g4.state_dict()
    assert torch.equal(sd4["GAB_UP.upsampling.3.weight"], g2.state_dict()["GAB_UP.upsampling.0.weight"])
    assert sd4["GAB_UP.upsampling.0.weight"].data_ptr() == sd4["GAB_UP.upsampling.3.weight"].data_ptr()


def test_chain_x2_x3_x4(emu, tmp_path):
    net = Recording(_args(tmp_path))
    finals, starts = {}, {}

    def stage_end(s, n):
        finals[s] = ({k: v.detach().clone() for k, v in n.generator.state_dict().items()},
                     {k: v.detach().clone() for k, v in n.discriminator.state_dict().items()})
        starts[s] = getattr(n, "started_from", None)
        n.started_from = None
        assert n.optimizer_G.step_count == 2 and n.optimizer_D.step_count == 2      # fresh Adam state per stage (:724-725)

    res = net.chain_train((2, 3, 4), on_stage_end=stage_end)
    assert sorted(res) == [2, 3, 4]
    assert all(len(r[0]) == 1 and all(map(lambda v: v == v, r[0])) for r in res.values())        # one finite epoch mean per stage
    assert starts[2] is None                                                                      # first stage: fresh init (:713-714)
    for prev, cur in ((2, 3), (3, 4)):
        g_start, d_start = starts[cur]
        g_prev, d_prev = finals[prev]
        for k, v in g_start.items():
            if k.startswith("GAB_UP.upsampling."):
                assert k not in g_prev or tuple(v.shape) != tuple(g_prev[k].shape)               # the two head families differ
            else:
                assert torch.equal(v, g_prev[k]), k
        assert all(torch.equal(d_start[k], d_prev[k]) for k in d_prev)                            # critic carried over entirely
    # every stage moved the weights it started from
    assert not torch.equal(finals[3][0]["conv1.0.weight"], starts[3][0]["conv1.0.weight"])
    assert (tmp_path / "x4" / "model" / "generator_param.pkl").exists()
    assert net.generator.GAB_UP.upsampling[0].weight.shape[0] == 256 and net.scale_factor == 4
