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
    ws[2].__dict__["_sr_pack"][(0, torch.float32, 0)] = (ops._ver(ws[2]), torch.empty(1, 64, 3), False)     # fp32 operand of an RGB-side thin layer
    plan = ops.PackPlan(ws)
    assert len(plan.entries) == 8 and plan.valid()
    assert [t[0] for t in plan.tables] == [torch.bfloat16, torch.float32]       # one batched launch per operand dtype
    total = 0
    for dt, t, n, blocks in plan.tables:
        assert t.shape == (n, 8) and t.dtype == torch.int64
        mine = [(w, key, out) for (w, key, out) in plan.entries if key[1] == dt]
        assert len(mine) == n == (7 if dt == torch.bfloat16 else 1)
        first = 0
        for row, (w, key, out) in zip(t.tolist(), mine):
            co, ci, kh, kw = w.shape
            assert row[0] == w.data_ptr() and row[1] == out.data_ptr()
            assert row[2:7] == [co, ci, kh * kw, key[0], key[2]]
            assert row[7] == first                              # every entry owns ceil(numel / 1024) consecutive blocks
            first += (w.numel() + 1023) // 1024
        assert blocks == first
        total += blocks
    assert plan.blocks == total
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
