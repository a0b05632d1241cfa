"""x9 tiled-inference throughput vs tiles per forward / tile size (bench.inference_bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for tile, tb in ((128, 32), (128, 64), (160, 32)):
    r = bench.inference_bench(2048, tile=tile, tile_batch=tb)
    print("tile %d  tiles/forward %d  -> %.1f Mpix/s (%.3f s)" % (tile, tb, r["value"], r["seconds"]), flush=True)
