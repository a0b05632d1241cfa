"""L2 write bandwidth vs store pattern (sr_debug_store_pattern, diagnostics build): one launch writes rows x row_bytes once."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sradsgan_b200 import _lib
lib = _lib.load()
rows, row_bytes = 46656 // 32 * 32, 512            # the RAB conv1 output: 46 656 pixels x 256 bf16 channels
buf = torch.empty(rows * row_bytes, dtype=torch.uint8, device="cuda")
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for grid in (148, 296):
    for pattern in (0, 1, 2, 3):
        ts = []
        for rep in range(6):
            big.zero_()                                # evict the output lines from L2
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.sr_debug_store_pattern(ctypes.c_void_p(buf.data_ptr()), rows, row_bytes, pattern, grid, None)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[1]
        print("grid %3d pattern %d: %7.1f us  %6.0f GB/s  (rc %d)" % (grid, pattern, t, rows * row_bytes / t / 1e3, rc), flush=True)
