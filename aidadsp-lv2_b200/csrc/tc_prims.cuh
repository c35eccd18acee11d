// tc_prims.cuh -- thin inline-PTX wrappers for the sm_100a tensor-core path: tcgen05.mma (kind::tf32) with
// shared-memory operand descriptors, TMEM allocation / loads, mbarrier completion and the proxy fences.
// Only what kernels_tc.cuh needs; no CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace aidax {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (one arrival per phase: the tcgen05.commit of the MMA-issuing thread)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar_addr, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity), "r"(0x989680u) /* suspend-time hint: park the warp instead of spinning on issue slots */
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar_addr, uint32_t parity)
{
    while (!mbar_try_wait(bar_addr, parity)) {}
}

__device__ __forceinline__ void prefetch_l2(const void* gptr)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(gptr)));
}

// ---- fences
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// named barrier over `nthreads` threads (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// producer side of the same barrier: signals arrival and goes on (the PTX producer/consumer pattern: the warps that only
// need to tell the MMA issuer "my operand rows are written" do not wait for each other)
__device__ __forceinline__ void bar_arrive(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMEM allocation (whole warp executes; column count a power of two >= 32)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}

// ---- operand descriptors
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): the operand is a grid of 8-row x 16-byte core
// matrices, each 128 contiguous bytes.  lbo = byte distance between core matrices adjacent along K,
// sbo = byte distance between core matrices adjacent along M/N.  Bits: [0,14) address>>4, [16,30) lbo>>4,
// [32,46) sbo>>4, [46,48) descriptor version (1 on sm_100), [61,64) swizzle mode (0 = none).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, M x N tile.
// Bits: [4,6) D format (1 = f32), [7,10) A format (2 = tf32), [10,13) B format, [15] A major, [16] B major,
// [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] . B[smem]^T, issued by ONE thread for the whole CTA
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}"
        :
        : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand in TENSOR MEMORY ("TS" form): A[m][k] (32-bit tf32 elements) sits at lane m, column a_tmem + k
// (cute::UMMA::tmem_frg for M = 128: datapath = row, one column per 32-bit element); a K step of 8 advances 8 columns.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> TMEM: 4 consecutive 32-bit columns of this thread's own lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
}
// this thread's tcgen05.st have completed (then fence::before_thread_sync + a barrier make them visible to the MMA issuer)
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]) : : "memory");
}

// arrive on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint32_t bar_addr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}

// ---- TMEM -> registers: 16 consecutive fp32 columns of this thread's own lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// wait for this thread's outstanding tcgen05.ld; the registers are threaded through as in/out operands so
// that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// ---- 3xTF32 operand split: x = hi + lo with hi = x truncated to tf32 (exactly what the tensor core would
// keep) and lo = x - hi (exact in fp32).  The tensor core truncates lo to its top 11 bits; adding half a tf32 ulp
// to lo's bit pattern first turns that truncation into round-to-nearest, which lowers the rms error of the
// product by ~40 % (tools/tc_gemm_probe.cu modes 0 vs 4, measured on B200: 3.2e-7 -> 2.0e-7 at K = 24).
// hi.B_hi + hi.B_lo + lo.B_hi then recovers the fp32 product to ~2^-22 relative.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo)
{
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
}
// Same with hi ROUNDED to the nearest tf32 value: lo is then signed and at most half a tf32 ulp, so the dropped
// lo.lo term of the 3-product scheme is half as large and has no systematic sign (with both operands truncated
// it always has the sign of the exact product).  One more integer add per element: used where the split is built
// once (the weights), while the per-step split of the recurrent vector stays truncating.
__device__ __forceinline__ void split_tf32_rn(float x, float& hi, float& lo)
{
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
}

// ---- packed fp32 pairs (sm_100 FMUL2 / FADD2 / FFMA2): two IEEE round-to-nearest operations per issue slot.
// The gate epilogue is bound by issue slots and the XU pipe, not by FMA-pipe throughput, so its fp32 arithmetic
// runs two hidden units per instruction.  Each half is bit-identical to the scalar mul.rn / add.rn / fma.rn.
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b)
{
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(r.x), "=f"(r.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

} // namespace tc
} // namespace aidax
