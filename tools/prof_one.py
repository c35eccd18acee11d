"""Scratch: run one configuration a few times (for ncu).  python tools/prof_one.py N H tile_mode [iters] [cell] [in]"""
import sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from quick_bench import run
a = [int(v) for v in sys.argv[1:]]
run(a[0], a[1], tile_mode=a[2], iters=a[3] if len(a) > 3 else 3, cell=a[4] if len(a) > 4 else 0, input_size=a[5] if len(a) > 5 else 1)
