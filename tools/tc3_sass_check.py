"""Scratch: static check of the warp-specialised kernels in an object file: registers used by each role must stay
below its setmaxnreg budget (a callee or a mis-allocated region would touch registers the warp no longer owns), and
local-memory spills should sit in the filter role only.   python tools/tc3_sass_check.py build/variants_h16.o"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
bad = 0
for fn in out.split("Function : ")[1:]:
    name = fn.split("\n", 1)[0].strip()
    if "chain_kernel_tc3" not in name:
        continue
    lines = [l for l in fn.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)]
    dealloc = next(i for i, l in enumerate(lines) if "USETMAXREG.DEALLOC" in l)
    alloc = next(i for i, l in enumerate(lines) if "USETMAXREG.TRY_ALLOC" in l)
    fregs = int(re.search(r"DEALLOC\.CTAPOOL (0x[0-9a-f]+)", lines[dealloc]).group(1), 16)
    rregs = int(re.search(r"TRY_ALLOC\.CTAPOOL \S+ (0x[0-9a-f]+)", lines[alloc]).group(1), 16)
    last_exit = max(i for i, l in enumerate(lines) if " EXIT" in l)
    def stats(lo, hi):
        regs = [int(m) for l in lines[lo:hi] for m in re.findall(r"\bR(\d+)\b", l)]
        spills = sum(1 for l in lines[lo:hi] if re.search(r"\b(STL|LDL)\b", l))
        return (max(regs) if regs else -1), spills, hi - lo
    pre, f, r, tail = stats(0, dealloc), stats(dealloc, alloc), stats(alloc, last_exit + 1), stats(last_exit + 1, len(lines))
    ok = f[0] < fregs and r[0] < rregs and pre[0] < min(fregs, rregs) + 64 and r[1] == 0
    # the subroutines behind the last EXIT (division slow path, mbarrier retry loops) are shared by both roles
    ok = ok and tail[0] < fregs
    bad += 0 if ok else 1
    print("%s\n   F budget %d: max R%d, %d spill instrs, %d instrs | R budget %d: max R%d, %d spill instrs, %d instrs | prologue max R%d | tail max R%d  %s" %
          (name[:110], fregs, f[0], f[1], f[2], rregs, r[0], r[1], r[2], pre[0], tail[0], "ok" if ok else "CHECK"))
sys.exit(bad)
