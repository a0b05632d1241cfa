"""GPU: chain training (BASELINE.json configs[2]) on the CUDA path — x2 -> x3 -> x4 stages of a small SRADSGAN, each a
complete `train()` on synthetic batches (graph-free eager steps), warm-started from the previous stage."""
import pytest
import torch

from test_chain_training_cpu import Recording, _args

pytestmark = pytest.mark.gpu


def test_chain_x2_x3_x4_on_gpu(tmp_path):
    from sradsgan_b200 import ops
    prev = ops.config.compute_dtype
    try:
        net = Recording(_args(tmp_path, precision="bf16", batch_size=2, crop_size=72, test_crop_size=72, hr_height=72, hr_width=72))
        seen = {}

        def stage_end(s, n):
            seen[s] = ({k: v.detach().float().cpu().clone() for k, v in n.generator.state_dict().items()},
                       getattr(n, "started_from", None))
            n.started_from = None
            assert n.optimizer_G.step_count == 2 and n.optimizer_D.step_count == 2

        res = net.chain_train((2, 3, 4), on_stage_end=stage_end)
        assert sorted(res) == [2, 3, 4]
        for s, (lg, ld) in res.items():
            assert len(lg) == 1 and lg[0] == lg[0] and ld[0] == ld[0], (s, lg, ld)          # finite epoch means
        for prev_s, cur in ((2, 3), (3, 4)):
            start = seen[cur][1][0]
            final_prev = seen[prev_s][0]
            carried = [k for k in start if not k.startswith("GAB_UP.upsampling.")]
            assert carried and all(torch.equal(start[k].float().cpu(), final_prev[k]) for k in carried)
        assert net.generator.GAB_UP.upsampling[0].weight.shape[0] == 256 and net.generator.GAB_UP.upsampling[0].weight.is_cuda
    finally:
        ops.config.compute_dtype = prev
