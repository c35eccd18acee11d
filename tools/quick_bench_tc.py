"""Scratch: tensor-core shape at several stream counts."""
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from quick_bench import run
if __name__ == "__main__":
    for n in (18944, 75776, 131072, 262144):
        run(n, 16, tile_mode=3)
    run(131072, 16, tile_mode=2)
    run(262144, 40, tile_mode=3)
    run(131072, 16, cell=1, input_size=3, tile_mode=3)
    run(131072, 12, tile_mode=3); run(131072, 8, tile_mode=3); run(131072, 24, tile_mode=3); run(131072, 32, tile_mode=3)
