"""Scratch: small-batch latency of the FP32 shapes (tile_mode 1 = lowest-latency shape, 0 = auto, 2 = throughput)."""
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from quick_bench import run
if __name__ == "__main__":
    for n in (1, 128, 1024, 2048, 4096, 8192, 16384):
        for tm in (1, 0):
            run(n, 16, tile_mode=tm, iters=50)
    run(1024, 12, tile_mode=1, iters=50); run(1024, 8, tile_mode=1, iters=50)
    run(1024, 16, cell=1, input_size=3, tile_mode=1, iters=50)
    run(8192, 16, cell=1, input_size=3, tile_mode=1, iters=50); run(8192, 16, cell=1, input_size=3, tile_mode=0, iters=50)
