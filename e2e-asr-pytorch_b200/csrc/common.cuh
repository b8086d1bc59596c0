// Shared device helpers for the sm_100a decode kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "../../include/e2e_asr_b200.h"

#define E2E_LN2F 0.693147180559945309417232121458176568f
#define E2E_FULL_MASK 0xffffffffu

namespace e2e {

// ---- host-side error plumbing (abi.cu) ---------------------------------------------------
int set_error(int code, const char *fmt, ...);
int check_launch(const char *what);
void count_launch(int n = 1);

// ---- log-add-exp ---------------------------------------------------------------------------
// numpy's fp32 logaddexp (the arithmetic src/ctc.py runs on):
//   x==y -> x+ln2 ; d=x-y ; d>0 -> x+log1p(exp(-d)) ; else y+log1p(exp(d))
// which is max(x,y) + log1p(exp(-|x-y|)) in both branches.
template <bool kFast>
__device__ __forceinline__ float logaddexp(float a, float b)
{
    const float m = fmaxf(a, b);
    const float d = -fabsf(a - b);
    float l;
    if (kFast) {
        // exp(d) in (0,1]; 1+e in (1,2]; two MUFU ops.  Absolute error ~2e-7, the same order as
        // expf+log1pf here because e <= 1 (see DESIGN.md "numerics").
        l = __logf(1.0f + __expf(d));
    } else {
        l = (d == 0.0f) ? E2E_LN2F : log1pf(expf(d));
    }
    return __fadd_rn(m, l);
}

// ---- warp reductions -----------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(E2E_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(E2E_FULL_MASK, v, o);
    return v;
}
// arg-max with ties resolved towards the LOWER index; idx == INT_MAX means "nothing".
__device__ __forceinline__ void warp_argmax(float &v, int &i)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(E2E_FULL_MASK, v, o);
        const int oi = __shfl_xor_sync(E2E_FULL_MASK, i, o);
        if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}

// ---- mbarrier + bulk async copy (TMA engine, SASS: UBLKCP / SYNCS) --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace e2e
