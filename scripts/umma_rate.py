"""tcgen05.mma issue-rate probe (sr_debug_umma_rate): cycles per instruction vs N, number of independent accumulators,
the number of consecutive instructions into the same accumulator, and the number of concurrently issuing warps."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sradsgan_b200 import _lib
lib = _lib.load()
out = torch.zeros(148 * 4, dtype=torch.int64, device="cuda")
print("N  issuers acc k_steps x4 -> SM cycles per MMA (wall / instructions issued by ALL issuers)   floor = 128*N/256")
grid = 148
for n in (64, 128, 256):
    for issuers, acc, ks, x4 in ((1, 1, 4, 0), (1, 1, 4, 1), (1, 1, 64, 1), (2, 1, 4, 0), (2, 1, 4, 1), (4, 1, 4, 0), (4, 1, 4, 1), (2, 2, 4, 1)):
        if n * acc * issuers > 512:
            continue
        iters = max(1, 2048 // (acc * ks))
        code = acc | (issuers << 8) | (x4 << 16)
        rc = lib.sr_debug_umma_rate(n, code, iters, ks, grid, ctypes.c_void_p(out.data_ptr()), None)
        torch.cuda.synchronize()
        per_issuer = iters * acc * ks
        wall = out[:grid * issuers].float().max().item()
        print("%4d %4d %5d %5d %3d      %8.1f   (floor %d)   rc=%d" % (n, issuers, acc, ks, x4, wall / (per_issuer * issuers), 128 * n // 256, rc), flush=True)
