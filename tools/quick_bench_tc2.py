import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from quick_bench import run
for n in (56832, 131072, 262144):
    run(n, 16, tile_mode=3)
run(131072, 16, tile_mode=3, eq=False)
