"""CPU: the shipped libaidax.so really contains what DESIGN.md 3.2c describes -- read off the SASS with cuobjdump (no GPU):
the default kernel of the bench workload is the warp-specialised tcgen05 kernel (UTCHMMA with the A operand in tensor
memory, STTM/LDTM, the MMA-completion barrier), the roles re-split their registers with setmaxnreg, and the recurrence
role runs without local-memory spills."""
import os
import re
import shutil
import subprocess

import pytest

import aidadsp_lv2_b200 as ax

# chain_kernel_tc3<CELL=LSTM, H=16, SPLIT=1, NFW=8, RREG=96, FREG=32, RES=2, MATH=fast, NTILES=2>
KERNEL = "_ZN5aidax16chain_kernel_tc3ILi0ELi16ELi1ELi8ELi96ELi32ELi2ELi0ELi2EEEvNS_12RenderParamsE"


@pytest.fixture(scope="module")
def sass(native_libs):
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", "-fun", KERNEL, ax.lib_path("libaidax.so")], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)]
    assert len(lines) > 1000, "kernel %s is not in libaidax.so" % KERNEL
    return lines


def test_tensor_core_path_is_tcgen05_with_the_operand_in_tensor_memory(sass):
    text = "\n".join(sass)
    assert "sm_100" in subprocess.run(["cuobjdump", "-lelf", ax.lib_path("libaidax.so")], capture_output=True, text=True).stdout
    mma = [l for l in sass if "UTCHMMA" in l]
    assert len(mma) >= 9                                  # 3 K steps x 3 products of the TF32 split, per unrolled copy of the time loop
    # TS form: A from tensor memory, B through a shared-memory descriptor, accumulator in tensor memory
    # (the SS form of generation 1 reads `UTCHMMA gdesc[..], gdesc[..], tmem[..]`)
    assert all(re.search(r"UTCHMMA\s+tmem\[\w+\],\s*gdesc\[\w+\],\s*tmem\[", l) for l in mma), mma[0]
    for op in ("STTM", "LDTM", "UTCBAR", "SYNCS.PHASECHK", "SYNCS.ARRIVE", "MUFU.EX2", "MUFU.RCP"):
        assert op in text, op


def test_roles_split_the_register_file(sass):
    dealloc = [l for l in sass if "USETMAXREG.DEALLOC" in l]
    alloc = [l for l in sass if "USETMAXREG.TRY_ALLOC" in l]
    assert len(dealloc) == 1 and len(alloc) == 1
    assert int(re.search(r"CTAPOOL\s+(0x[0-9a-f]+)", dealloc[0]).group(1), 16) == 32   # filter warps
    assert int(re.search(r"(0x[0-9a-f]+)\s*;", alloc[0]).group(1), 16) == 96          # recurrence warps
    # nothing in the filter role may be a call into a separately allocated function: only the division slow path and the
    # barrier retry stubs sit behind the last EXIT, and they stay inside the filter role's 32 registers
    last_exit = max(i for i, l in enumerate(sass) if re.search(r"\bEXIT\b", l))
    tail_regs = [int(m) for l in sass[last_exit + 1:] for m in re.findall(r"\bR(\d+)\b", l)]
    assert max(tail_regs) < 32


def test_recurrence_time_loop_has_no_spills(sass):
    # the time loop = the region between the first and the last MUFU.EX2 of the longest MUFU-dense stretch after the
    # register re-allocation; a spill there would be an STL/LDL
    alloc = next(i for i, l in enumerate(sass) if "USETMAXREG.TRY_ALLOC" in l)
    last_exit = max(i for i, l in enumerate(sass) if re.search(r"\bEXIT\b", l))
    region = sass[alloc:last_exit + 1]
    mufu = [i for i, l in enumerate(region) if "MUFU" in l]
    assert len(mufu) >= 100
    spills = [l for l in region[mufu[0]:mufu[-1] + 1] if re.search(r"\b(STL|LDL)\b", l)]
    assert not spills, spills[:3]
