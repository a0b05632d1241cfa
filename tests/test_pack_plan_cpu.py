"""CPU: host logic of ops.PackPlan (the table handed to sr_pack_weights_batched) and of the packed-operand cache stamps —
no kernel runs here; the launch itself is covered by tests/test_gpu_conv_kernels.py::test_batched_weight_packing_matches_single_packs."""
import torch

from sradsgan_b200 import ops


def _param(shape):
    return torch.nn.Parameter(torch.randn(*shape))


def test_plan_table_layout_and_validity():
    ws = [_param((256, 64, 3, 3)), _param((64, 256, 3, 3)), _param((64, 3, 1, 1)), _param((7,))]      # the 1-D bias is ignored
    for w in ws[:3]:
        co, ci, kh, kw = w.shape
        w.__dict__["_sr_pack"] = {(0, torch.bfloat16, 0): (ops._ver(w), torch.empty(kh * kw, co, ci, dtype=torch.bfloat16), False),
                                  (1, torch.bfloat16, 0): (ops._ver(w), torch.empty(kh * kw, ci, co, dtype=torch.bfloat16), False)}
    ws[0].__dict__["_sr_pack"][(0, torch.bfloat16, 2)] = (ops._ver(ws[0]), torch.empty(9, 256, 64, dtype=torch.bfloat16), False)
    ws[1].__dict__["_sr_pack"][(0, torch.float32, 0)] = (ops._ver(ws[1]), torch.empty(9, 64, 256), False)     # other dtype: not in this plan
    plan = ops.PackPlan(ws)
    assert plan.dtype == torch.bfloat16 and len(plan.entries) == 7 and plan.valid()
    t = plan.table
    assert t.shape == (7, 8) and t.dtype == torch.int64
    first = 0
    for row, (w, key, out) in zip(t.tolist(), plan.entries):
        co, ci, kh, kw = w.shape
        assert row[0] == w.data_ptr() and row[1] == out.data_ptr()
        assert row[2:7] == [co, ci, kh * kw, key[0], key[2]]
        assert row[7] == first                                  # every entry owns ceil(numel / 1024) consecutive blocks
        first += (w.numel() + 1023) // 1024
    assert plan.blocks == first
    ws[0].data = ws[0].data.clone()                             # master moved (e.g. a new flat buffer): the table is stale
    assert not plan.valid()


def test_cache_stamp_tracks_optimizer_generation_and_inplace_updates():
    w = _param((8, 4, 3, 3))
    v0 = ops._ver(w)
    w._sr_gen = ops.next_generation()                           # what FlatAdam.step does for its own parameters
    v1 = ops._ver(w)
    assert v1 != v0
    with torch.no_grad():
        w.mul_(2.0)                                             # in-place edit (load_state_dict copies in place)
    assert ops._ver(w) != v1
    g0 = ops.next_generation()
    assert ops.next_generation() == g0 + 1                      # stamps never repeat, also across optimizer objects
