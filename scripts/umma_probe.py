"""Runs the tcgen05 operand-descriptor probe (sr_debug_umma_shift) over shifts / group strides / base_offset values."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sradsgan_b200 import _lib

lib = _lib.load()
rows = 512
torch.manual_seed(0)
A = torch.randn(rows, 64, device="cuda").to(torch.bfloat16)
Bm = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
out = torch.empty(128, 64, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
for sbo in (1024, 1280, 2304):
    for shift in (0, 1, 2, 3, 5, 8, 9, 19):
        res = []
        for bo_mode in ("0", "shift&7"):
            bo = 0 if bo_mode == "0" else (shift & 7)
            out.zero_()
            rc = lib.sr_debug_umma_shift(p(A), rows, p(Bm), shift, sbo, bo, p(out), None)
            torch.cuda.synchronize()
            idx = torch.tensor([shift + (r // 8) * (sbo // 128) + r % 8 for r in range(128)], device="cuda")
            ref = A[idx].float() @ Bm.float().t()
            err = ((out - ref).norm() / ref.norm()).item()
            res.append("base_offset=%s: rc=%d rel.err=%.3e" % (bo_mode, rc, err))
        print("sbo=%4d shift=%2d | %s" % (sbo, shift, " | ".join(res)), flush=True)
