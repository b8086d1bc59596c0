// One-token LSTM step helpers for the batched RNNLM / speller step.
// The recurrent GEMMs stay library calls (bf16 tensor-core GEMMs on a 3-way split of the fp32
// operands, see stepper.py); these two kernels replace the ~25 element-wise launches per layer
// that used to surround them and remove the per-step state permutation:
//   lstm_split_rows_kernel : gathers fp32 rows through the parents' row index and writes their
//                            exact 3-piece bf16 split straight into the GEMM's A operand
//   lstm_cell_kernel       : gates (+ bias, + layer-0 input table) -> (c', h') in fp32 with the
//                            reference's operation order (nn.LSTM cell: src/lm.py:31, src/asr.py:262),
//                            reading c through the parents' row index, and writing the bf16 split
//                            of h' into the NEXT layer's A operand
// Both are pure HBM streams: one pass over their inputs and outputs, 16-byte accesses.
#include "common.cuh"
#include <cuda_bf16.h>

namespace e2e {

// x == a1 + a2 + a3 exactly (barring underflow): a1 = bf16(x), a2 = bf16(x - a1), a3 = bf16(x - a1 - a2)
__device__ __forceinline__ void split3(float x, __nv_bfloat16 &a1, __nv_bfloat16 &a2, __nv_bfloat16 &a3)
{
    a1 = __float2bfloat16_rn(x);
    const float r1 = __fsub_rn(x, __bfloat162float(a1));
    a2 = __float2bfloat16_rn(r1);
    a3 = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(a2)));
}

struct alignas(8) Bf16x4 { __nv_bfloat16 v[4]; };

// dst[r][p*K + off + c] = piece_p(src[row(r)][c]),  p = 0,1,2,  c < w;  row(r) = idx ? idx[r] : r
template <bool kVec>
__global__ void __launch_bounds__(256)
lstm_split_rows_kernel(const float *__restrict__ src, long long src_pitch, const long long *__restrict__ idx, int n, int w,
                       __nv_bfloat16 *__restrict__ dst, long long dst_pitch, int K, int off)
{
    const int r = blockIdx.y;
    const long long row = idx ? idx[r] : r;
    const float *s = src + row * src_pitch;
    __nv_bfloat16 *d = dst + (long long)r * dst_pitch + off;
    if (kVec) {
        for (int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4; c < w; c += gridDim.x * blockDim.x * 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(s + c));
            Bf16x4 p0, p1, p2;
            split3(v.x, p0.v[0], p1.v[0], p2.v[0]);
            split3(v.y, p0.v[1], p1.v[1], p2.v[1]);
            split3(v.z, p0.v[2], p1.v[2], p2.v[2]);
            split3(v.w, p0.v[3], p1.v[3], p2.v[3]);
            *reinterpret_cast<Bf16x4 *>(d + c) = p0;
            *reinterpret_cast<Bf16x4 *>(d + K + c) = p1;
            *reinterpret_cast<Bf16x4 *>(d + 2 * K + c) = p2;
        }
    } else {
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < w; c += gridDim.x * blockDim.x) {
            __nv_bfloat16 a1, a2, a3;
            split3(__ldg(s + c), a1, a2, a3);
            d[c] = a1; d[K + c] = a2; d[2 * K + c] = a3;
        }
    }
}

// the same for up to kMaxSplitSrc sources in one launch (blockIdx.z = source): all layers of an LSTM stack at once
constexpr int kMaxSplitSrc = 8;
struct SplitMulti {
    const float *src[kMaxSplitSrc]; long long src_pitch[kMaxSplitSrc]; int w[kMaxSplitSrc];
    __nv_bfloat16 *dst[kMaxSplitSrc]; long long dst_pitch[kMaxSplitSrc]; int K[kMaxSplitSrc], off[kMaxSplitSrc];
    const long long *idx; int n;
};

__global__ void __launch_bounds__(256)
lstm_split_rows_multi_kernel(const SplitMulti p)
{
    const int z = blockIdx.z, r = blockIdx.y;
    const long long row = p.idx ? p.idx[r] : r;
    const float *s = p.src[z] + row * p.src_pitch[z];
    __nv_bfloat16 *d = p.dst[z] + (long long)r * p.dst_pitch[z] + p.off[z];
    const int w = p.w[z], K = p.K[z];
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4; c < w; c += gridDim.x * blockDim.x * 4) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(s + c));
        Bf16x4 p0, p1, p2;
        split3(v.x, p0.v[0], p1.v[0], p2.v[0]);
        split3(v.y, p0.v[1], p1.v[1], p2.v[1]);
        split3(v.z, p0.v[2], p1.v[2], p2.v[2]);
        split3(v.w, p0.v[3], p1.v[3], p2.v[3]);
        *reinterpret_cast<Bf16x4 *>(d + c) = p0;
        *reinterpret_cast<Bf16x4 *>(d + K + c) = p1;
        *reinterpret_cast<Bf16x4 *>(d + 2 * K + c) = p2;
    }
}

__device__ __forceinline__ float sigmoid_f32(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

struct CellParams {
    const float *gates; long long gates_pitch; const float *bias; const float *table; const long long *tok;
    const float *c_prev; const long long *idx; int n, D;
    float *c_new, *h_new; __nv_bfloat16 *a_next; long long a_pitch; int K_next, off_next;
};

__device__ __forceinline__ void lstm_unit(const float (&g)[4], const float (&b)[4], const float *tb, float c, float &c2, float &h2)
{
    float z[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        z[q] = __fadd_rn(g[q], b[q]);                       // y + (b_ih + b_hh)
        if (tb) z[q] = __fadd_rn(tb[q], z[q]);              // + layer-0 input projection
    }
    c2 = __fadd_rn(__fmul_rn(sigmoid_f32(z[1]), c), __fmul_rn(sigmoid_f32(z[0]), tanhf(z[2])));
    h2 = __fmul_rn(sigmoid_f32(z[3]), tanhf(c2));
}

// one thread per (row, kW consecutive hidden units); gate order i, f, g, o (torch.nn.LSTM)
template <int kW>
__global__ void __launch_bounds__(256)
lstm_cell_kernel(const CellParams p)
{
    const int r = blockIdx.y;
    const float *g = p.gates + (long long)r * p.gates_pitch;
    const float *tb = p.table ? p.table + p.tok[r] * (long long)(4 * p.D) : nullptr;
    const long long prow = p.idx ? p.idx[r] : r;
    const float *cp = p.c_prev + prow * p.D;
    for (int d = (blockIdx.x * blockDim.x + threadIdx.x) * kW; d < p.D; d += gridDim.x * blockDim.x * kW) {
        float gv[4][kW], bv[4][kW], tv[4][kW], cv[kW];
        if (kW == 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(g + q * p.D + d));
                const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + q * p.D + d));
                gv[q][0] = a.x; gv[q][1] = a.y; gv[q][2] = a.z; gv[q][3] = a.w;
                bv[q][0] = b.x; bv[q][1] = b.y; bv[q][2] = b.z; bv[q][3] = b.w;
                if (tb) {
                    const float4 t = __ldg(reinterpret_cast<const float4 *>(tb + q * p.D + d));
                    tv[q][0] = t.x; tv[q][1] = t.y; tv[q][2] = t.z; tv[q][3] = t.w;
                }
            }
            const float4 c = __ldg(reinterpret_cast<const float4 *>(cp + d));
            cv[0] = c.x; cv[1] = c.y; cv[2] = c.z; cv[3] = c.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                gv[q][0] = __ldg(g + q * p.D + d);
                bv[q][0] = __ldg(p.bias + q * p.D + d);
                if (tb) tv[q][0] = __ldg(tb + q * p.D + d);
            }
            cv[0] = __ldg(cp + d);
        }
        float c2[kW], h2[kW];
#pragma unroll
        for (int i = 0; i < kW; ++i) {
            const float gi[4] = {gv[0][i], gv[1][i], gv[2][i], gv[3][i]};
            const float bi[4] = {bv[0][i], bv[1][i], bv[2][i], bv[3][i]};
            const float ti[4] = {tb ? tv[0][i] : 0.0f, tb ? tv[1][i] : 0.0f, tb ? tv[2][i] : 0.0f, tb ? tv[3][i] : 0.0f};
            lstm_unit(gi, bi, tb ? ti : nullptr, cv[i], c2[i], h2[i]);
        }
        float *co = p.c_new + (long long)r * p.D + d, *ho = p.h_new + (long long)r * p.D + d;
        if (kW == 4) {
            *reinterpret_cast<float4 *>(co) = make_float4(c2[0], c2[1], c2[2], c2[3]);
            *reinterpret_cast<float4 *>(ho) = make_float4(h2[0], h2[1], h2[2], h2[3]);
        } else {
            co[0] = c2[0]; ho[0] = h2[0];
        }
        if (p.a_next) {
            __nv_bfloat16 *o = p.a_next + (long long)r * p.a_pitch + p.off_next + d;
            if (kW == 4) {
                Bf16x4 p0, p1, p2;
#pragma unroll
                for (int i = 0; i < 4; ++i) split3(h2[i], p0.v[i], p1.v[i], p2.v[i]);
                *reinterpret_cast<Bf16x4 *>(o) = p0;
                *reinterpret_cast<Bf16x4 *>(o + p.K_next) = p1;
                *reinterpret_cast<Bf16x4 *>(o + 2 * p.K_next) = p2;
            } else {
                __nv_bfloat16 a1, a2, a3;
                split3(h2[0], a1, a2, a3);
                o[0] = a1; o[p.K_next] = a2; o[2 * p.K_next] = a3;
            }
        }
    }
}

}  // namespace e2e

extern "C" int e2e_lstm_split_rows(const float *src, long long src_pitch, const long long *row_idx, int n, int w,
                                   void *dst_bf16, long long dst_pitch, int K, int off, void *stream)
{
    using namespace e2e;
    if (!src || !dst_bf16) return set_error(E2E_ERR_ARG, "e2e_lstm_split_rows: null pointer");
    if (n <= 0 || w <= 0 || K <= 0 || off < 0 || off + w > K || dst_pitch < 3LL * K || src_pitch < w)
        return set_error(E2E_ERR_ARG, "e2e_lstm_split_rows: bad size");
    if (n > 65535 * 16) return set_error(E2E_ERR_UNSUPPORTED, "e2e_lstm_split_rows: too many rows");
    const bool vec = (w % 4 == 0) && (src_pitch % 4 == 0) && (dst_pitch % 4 == 0) && (K % 4 == 0) && (off % 4 == 0) &&
                     !(reinterpret_cast<uintptr_t>(src) & 15) && !(reinterpret_cast<uintptr_t>(dst_bf16) & 7);
    const int per_thread = vec ? 4 : 1;
    const int bx = (w + 256 * per_thread - 1) / (256 * per_thread);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    __nv_bfloat16 *dst = static_cast<__nv_bfloat16 *>(dst_bf16);
    for (int r0 = 0; r0 < n; r0 += 65535) {          // gridDim.y limit
        const int rows = n - r0 < 65535 ? n - r0 : 65535;
        const dim3 grid(bx, rows);
        if (vec)
            lstm_split_rows_kernel<true><<<grid, 256, 0, st>>>(src + (row_idx ? 0 : (long long)r0 * src_pitch), src_pitch,
                                                               row_idx ? row_idx + r0 : nullptr, rows, w,
                                                               dst + (long long)r0 * dst_pitch, dst_pitch, K, off);
        else
            lstm_split_rows_kernel<false><<<grid, 256, 0, st>>>(src + (row_idx ? 0 : (long long)r0 * src_pitch), src_pitch,
                                                                row_idx ? row_idx + r0 : nullptr, rows, w,
                                                                dst + (long long)r0 * dst_pitch, dst_pitch, K, off);
        count_launch();
    }
    return check_launch("e2e_lstm_split_rows");
}

extern "C" int e2e_lstm_split_rows_multi(int n_src, const float *const *srcs, const long long *src_pitches, const int *widths,
                                         void *const *dsts_bf16, const long long *dst_pitches, const int *Ks, const int *offs,
                                         const long long *row_idx, int n, void *stream)
{
    using namespace e2e;
    if (!srcs || !src_pitches || !widths || !dsts_bf16 || !dst_pitches || !Ks || !offs)
        return set_error(E2E_ERR_ARG, "e2e_lstm_split_rows_multi: null pointer");
    if (n_src <= 0 || n_src > kMaxSplitSrc || n <= 0) return set_error(E2E_ERR_ARG, "e2e_lstm_split_rows_multi: bad size");
    SplitMulti p;
    int wmax = 0;
    for (int i = 0; i < n_src; ++i) {
        if (!srcs[i] || !dsts_bf16[i]) return set_error(E2E_ERR_ARG, "e2e_lstm_split_rows_multi: null pointer");
        const bool vec = (widths[i] % 4 == 0) && (src_pitches[i] % 4 == 0) && (dst_pitches[i] % 4 == 0) && (Ks[i] % 4 == 0) &&
                         (offs[i] % 4 == 0) && !(reinterpret_cast<uintptr_t>(srcs[i]) & 15) && !(reinterpret_cast<uintptr_t>(dsts_bf16[i]) & 7);
        if (!vec) return set_error(E2E_ERR_UNSUPPORTED, "e2e_lstm_split_rows_multi: source %d is not 16-byte vectorisable", i);
        if (widths[i] <= 0 || offs[i] < 0 || offs[i] + widths[i] > Ks[i] || dst_pitches[i] < 3LL * Ks[i] || src_pitches[i] < widths[i])
            return set_error(E2E_ERR_ARG, "e2e_lstm_split_rows_multi: bad geometry of source %d", i);
        p.src[i] = srcs[i]; p.src_pitch[i] = src_pitches[i]; p.w[i] = widths[i];
        p.dst[i] = static_cast<__nv_bfloat16 *>(dsts_bf16[i]); p.dst_pitch[i] = dst_pitches[i]; p.K[i] = Ks[i]; p.off[i] = offs[i];
        wmax = widths[i] > wmax ? widths[i] : wmax;
    }
    const int bx = (wmax + 1023) / 1024;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int r0 = 0; r0 < n; r0 += 65535) {
        const int rows = n - r0 < 65535 ? n - r0 : 65535;
        SplitMulti q = p;
        q.n = rows;
        q.idx = row_idx ? row_idx + r0 : nullptr;
        for (int i = 0; i < n_src; ++i) {
            if (!row_idx) q.src[i] = p.src[i] + (long long)r0 * p.src_pitch[i];
            q.dst[i] = p.dst[i] + (long long)r0 * p.dst_pitch[i];
        }
        lstm_split_rows_multi_kernel<<<dim3(bx, rows, n_src), 256, 0, st>>>(q);
        count_launch();
    }
    return check_launch("e2e_lstm_split_rows_multi");
}

extern "C" int e2e_lstm_cell(const float *gates, long long gates_pitch, const float *bias, const float *table, const long long *tok,
                             const float *c_prev, const long long *row_idx, int n, int D,
                             float *c_new, float *h_new, void *a_next_bf16, long long a_pitch, int K_next, int off_next,
                             void *stream)
{
    using namespace e2e;
    if (!gates || !bias || !c_prev || !c_new || !h_new || (table && !tok))
        return set_error(E2E_ERR_ARG, "e2e_lstm_cell: null pointer");
    if (n <= 0 || D <= 0 || gates_pitch < 4LL * D) return set_error(E2E_ERR_ARG, "e2e_lstm_cell: bad size");
    if (a_next_bf16 && (K_next <= 0 || off_next < 0 || off_next + D > K_next || a_pitch < 3LL * K_next))
        return set_error(E2E_ERR_ARG, "e2e_lstm_cell: bad next-layer operand geometry");
    CellParams p;
    p.gates = gates; p.gates_pitch = gates_pitch; p.bias = bias; p.table = table; p.tok = tok;
    p.c_prev = c_prev; p.idx = row_idx; p.D = D; p.a_pitch = a_pitch; p.K_next = K_next; p.off_next = off_next;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = (D % 4 == 0) && (gates_pitch % 4 == 0) && !(reinterpret_cast<uintptr_t>(gates) & 15) &&
                     !(reinterpret_cast<uintptr_t>(bias) & 15) && !(reinterpret_cast<uintptr_t>(table) & 15) &&
                     !(reinterpret_cast<uintptr_t>(c_prev) & 15) && !(reinterpret_cast<uintptr_t>(c_new) & 15) &&
                     !(reinterpret_cast<uintptr_t>(h_new) & 15) &&
                     (!a_next_bf16 || ((a_pitch % 4 == 0) && (K_next % 4 == 0) && (off_next % 4 == 0) && !(reinterpret_cast<uintptr_t>(a_next_bf16) & 7)));
    const int per_thread = vec ? 4 : 1;
    const int threads = D / per_thread >= 256 ? 256 : ((D / per_thread + 31) / 32 * 32);
    const int bx = (D + threads * per_thread - 1) / (threads * per_thread);
    for (int r0 = 0; r0 < n; r0 += 65535) {
        const int rows = n - r0 < 65535 ? n - r0 : 65535;
        p.n = rows;
        p.gates = gates + (long long)r0 * gates_pitch;
        p.tok = tok ? tok + r0 : nullptr;
        p.idx = row_idx ? row_idx + r0 : nullptr;
        p.c_prev = row_idx ? c_prev : c_prev + (long long)r0 * D;
        p.c_new = c_new + (long long)r0 * D;
        p.h_new = h_new + (long long)r0 * D;
        p.a_next = a_next_bf16 ? static_cast<__nv_bfloat16 *>(a_next_bf16) + (long long)r0 * a_pitch : nullptr;
        if (vec) lstm_cell_kernel<4><<<dim3(bx, rows), threads, 0, st>>>(p);
        else lstm_cell_kernel<1><<<dim3(bx, rows), threads, 0, st>>>(p);
        count_launch();
    }
    return check_launch("e2e_lstm_cell");
}
