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
// which is max(x,y) + softplus(-|x-y|) in both branches, softplus(d) = log1p(exp(d)).
// Three interchangeable softplus evaluators:
//   kMathLut   (default) piecewise-cubic table on [-4,0] in shared memory + MUFU.EX2 series tail;
//              within one fp32 rounding of the exact value, bit-equal to glibc's log1pf(expf(d))
//              for ~3/4 of inputs, branch free, ~23 instructions;
//   kMathMufu  MUFU.EX2 + MUFU.LG2 (absolute error ~3e-7), ~8 instructions;
//   kMathLibm  CUDA expf + log1pf (branchy, ~55 instructions) — kept as a cross-check;
//   kMathPoly  MUFU.EX2 + a degree-8 polynomial in e = exp(-|x-y|): softplus = e*P(e) (tools/gen_softplus_poly.py),
//              no table, no tail case, ~14 instructions (kMathPolyEstrin: the same polynomial evaluated pairwise,
//              3 more instructions, half the dependent FMA chain).  The polynomial is good to 3.3e-8; in emulated
//              fp32 the whole recursion stays within 1 ulp of the oracle (tools/emulate_prefix_math.py).  Opt-in
//              until measured on the GPU.
enum { kMathLut = 0, kMathMufu = 1, kMathLibm = 2, kMathPoly = 3, kMathPolyEstrin = 4 };
#include "softplus_poly.inc"

constexpr int kLutNodes = 65;     // d in [-4, 0], spacing 1/16
constexpr int kLutCopies = 8;     // one copy per lane of a quarter warp: LDS.128 never bank-conflicts
__device__ const float4 kSoftplusLut[kLutNodes] = {
#include "softplus_lut.inc"
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Copies the table into shared memory, kLutCopies interleaved replicas: entry i, copy k at [i*8+k].
__device__ __forceinline__ void softplus_lut_to_smem(float4 *dst, int tid, int nt)
{
    for (int i = tid; i < kLutNodes * kLutCopies; i += nt) dst[i] = kSoftplusLut[i / kLutCopies];
}

// ad = |a-b| >= 0; returns softplus(-ad).  lut points at this lane's replica (base + (lane & 7)).
// tm = -16*min(ad,4) + (1.5*2^23 + 64) has the node index in its low mantissa bits:
// bits(tm) = 0x4B400000 + idx, idx in [0,64].  The table entry sits idx*128 bytes after the replica
// base, so its shared-memory address is (bits(tm) << 7) + (base - (0x4B400000 << 7)) in 32-bit
// wrap-around arithmetic: one LEA instead of mask + shift + add.
__device__ __forceinline__ float softplus_lut(float ad, uint32_t lut_adj)
{
    const float kMagic = 12582912.0f + 64.0f;           // 1.5*2^23 + index bias
    const float tc = fminf(ad, 4.0f);                    // clamped to the table, [-4, 0] in node units of 1/16
    const float tm = fmaf(tc, -16.0f, kMagic);           // low mantissa bits = round(-16 tc) + 64
    const float f = fmaf(tc, -16.0f, __fsub_rn(kMagic, tm)); // exact fraction in [-0.5, 0.5]
    float4 c;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
        : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w)
        : "r"((static_cast<uint32_t>(__float_as_int(tm)) << 7) + lut_adj));
    const float g = fmaf(f, fmaf(f, fmaf(f, c.w, c.z), c.y), c.x);
    // d < -4: e = exp(d) <= 0.0184, log1p(e) = e(1 - e/2 + e^2/3 - e^3/4) to 4e-10
    const float e = ex2_approx(ad * -1.4426950408889634f);
    const float q = fmaf(e, fmaf(e, fmaf(e, -0.25f, 0.333333343f), -0.5f), 1.0f);
    return ad > 4.0f ? e * q : g;
}
// The address operand of softplus_lut for a thread whose replica starts at `lut`.
__device__ __forceinline__ uint32_t softplus_lut_adj(const float4 *lut)
{
    return smem_u32(lut) - (0x4B400000u << 7);
}

// ad = |a-b| >= 0; softplus(-ad) = log1p(e) = e*P(e) with e = exp(-ad)
template <bool kEstrin>
__device__ __forceinline__ float softplus_poly(float ad)
{
    const float e = ex2_approx(ad * -1.4426950408889634f);
    float p;
    if (!kEstrin) {
        p = fmaf(kSoftplusPolyC8, e, kSoftplusPolyC7);
        p = fmaf(p, e, kSoftplusPolyC6);
        p = fmaf(p, e, kSoftplusPolyC5);
        p = fmaf(p, e, kSoftplusPolyC4);
        p = fmaf(p, e, kSoftplusPolyC3);
        p = fmaf(p, e, kSoftplusPolyC2);
        p = fmaf(p, e, kSoftplusPolyC1);
        p = fmaf(p, e, kSoftplusPolyC0);
    } else {
        const float e2 = e * e, e4 = e2 * e2;
        const float p01 = fmaf(kSoftplusPolyC1, e, kSoftplusPolyC0), p23 = fmaf(kSoftplusPolyC3, e, kSoftplusPolyC2);
        const float p45 = fmaf(kSoftplusPolyC5, e, kSoftplusPolyC4), p67 = fmaf(kSoftplusPolyC7, e, kSoftplusPolyC6);
        const float lo = fmaf(p23, e2, p01), hi = fmaf(p67, e2, p45);
        p = fmaf(fmaf(kSoftplusPolyC8, e4, hi), e4, lo);
    }
    return e * p;
}

template <int kMath>
__device__ __forceinline__ float logaddexp(float a, float b, uint32_t lut)
{
    const float m = fmaxf(a, b);
    const float ad = fabsf(a - b);
    float l;
    if (kMath == kMathLut) {
        l = softplus_lut(ad, lut);
    } else if (kMath == kMathPoly || kMath == kMathPolyEstrin) {
        l = softplus_poly<kMath == kMathPolyEstrin>(ad);
    } else if (kMath == kMathMufu) {
        l = lg2_approx(1.0f + ex2_approx(ad * -1.4426950408889634f)) * E2E_LN2F;
    } else {
        l = (ad == 0.0f) ? E2E_LN2F : log1pf(expf(-ad));
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
