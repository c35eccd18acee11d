// variants.h -- table of compiled chain_kernel instantiations (one translation unit per hidden size).
#pragma once
#include <stddef.h>
#include "layout.h"

namespace aidax {

struct KernelVariant {
    int cell;   // 0 LSTM, 1 GRU, 2 none
    int H, L, S, NW;
    int ws;     // streams per warp
    const void* fn[2]; // [math_mode] kernel entry (for cudaFuncSetAttribute / cudaLaunchKernel)
    int (*warp_smem_floats)(int ts, int n_param_tiles);
    int kind;             // 0 = FP32 register-tiled recurrence (kernels.cuh), 1 = tcgen05 tensor-core recurrence (kernels_tc.cuh),
                          // 2 = FP32 warp-per-stream latency shape (kernels_w.cuh)
    int fixed_smem_floats; // shared memory that does not scale with the sub-block length (weights / MMA operands)
    int ctas_per_sm;      // residency the shared-memory budget is sized for
    int cs;               // streams per CTA if it is not ws * NW (kind 3: always 256), else 0
};

// defined in variants.cu, compiled once per -DAIDAX_VARIANT_H=<H>
#define AIDAX_DECLARE_VARIANTS(Hval) extern const KernelVariant kVariantsH##Hval[]; extern const int kNumVariantsH##Hval;
AIDAX_DECLARE_VARIANTS(0)  // chain-only (no model)
AIDAX_DECLARE_VARIANTS(8)
AIDAX_DECLARE_VARIANTS(12)
AIDAX_DECLARE_VARIANTS(16)
AIDAX_DECLARE_VARIANTS(20)
AIDAX_DECLARE_VARIANTS(24)
AIDAX_DECLARE_VARIANTS(32)
AIDAX_DECLARE_VARIANTS(40)
AIDAX_DECLARE_VARIANTS(64)
AIDAX_DECLARE_VARIANTS(80)

} // namespace aidax
