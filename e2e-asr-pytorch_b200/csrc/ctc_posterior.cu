// Kernel (1): fused ReLU + log-softmax of the CTC head, written frame-major [Tmax][U][Vp]
// with 16-byte vector stores, and the blank running sum that seeds the prefix states.
// Replaces src/decode.py:94-95 (minus the cuBLAS Linear) and src/ctc.py:19-27.
#include "common.cuh"

namespace e2e {

constexpr int kRowWarps = 8;   // rows (one per warp) per CTA

// Narrow rows (Vp <= 128, e.g. the 31-token character vocabulary): a row is owned by a group of G = Vp/4 lanes (rounded up
// to a power of two), each lane holding 4 columns in registers, so a warp carries 32/G rows (4 rows at Vp = 32): the
// max / sum-exp reductions are log2(G) shuffles, and the warp's rows leave as one contiguous run of 16-byte stores
// (consecutive rows t*U+u are consecutive in the frame-major tensor).
template <int G>
__global__ void __launch_bounds__(256)
ctc_log_softmax_narrow_kernel(const float *__restrict__ logits, int U, int Tmax, int V, const int *__restrict__ enc_len,
                              int apply_relu, float *__restrict__ x, int Vp)
{
    constexpr int kRowsPerWarp = 32 / G;
    const int lane = threadIdx.x & 31, sub = lane / G, gl = lane % G;
    const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kRowsPerWarp + sub;   // = t*U + u
    const bool valid_row = row < (long long)Tmax * U;
    const int t = valid_row ? (int)(row / U) : 0, u = valid_row ? (int)(row % U) : 0;
    const int tu = enc_len ? enc_len[u] : Tmax;
    const bool live = valid_row && t < tu;
    const int v0 = gl * 4;
    const float *__restrict__ in = logits + ((long long)u * Tmax + t) * V;
    float a[4];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float val = -INFINITY;
        if (live && v0 + k < V) {
            val = __ldg(in + v0 + k);
            if (apply_relu) val = fmaxf(val, 0.0f);
        }
        a[k] = val;
        m = fmaxf(m, val);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(E2E_FULL_MASK, m, o));
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (a[k] == -INFINITY) ? 0.0f : expf(a[k] - m);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(E2E_FULL_MASK, s, o);
    const float ls = logf(s);
    if (valid_row && v0 < Vp) {
        float o4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) o4[k] = (live && v0 + k < V) ? (a[k] - m) - ls : E2E_CTC_LOGZERO;
        reinterpret_cast<float4 *>(x + row * Vp)[gl] = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
}

// One warp per output row (t,u); lane l owns the 4-column groups l, l+32, ... of the row (rows of any width; the fallback
// for rows too wide for shared-memory staging).
template <bool kCached>
__global__ void __launch_bounds__(kRowWarps * 32)
ctc_log_softmax_kernel(const float *__restrict__ logits, int U, int Tmax, int V, const int *__restrict__ enc_len,
                       int apply_relu, float *__restrict__ x, int Vp)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kRowWarps + (threadIdx.x >> 5);   // = t*U + u
    if (row >= (long long)Tmax * U) return;
    const int t = (int)(row / U), u = (int)(row % U);
    float4 *__restrict__ out = reinterpret_cast<float4 *>(x + row * Vp);
    const int groups = Vp >> 2;
    const int tu = enc_len ? enc_len[u] : Tmax;
    if (t >= tu) {
        const float4 z = make_float4(E2E_CTC_LOGZERO, E2E_CTC_LOGZERO, E2E_CTC_LOGZERO, E2E_CTC_LOGZERO);
        for (int g = lane; g < groups; g += 32) out[g] = z;
        return;
    }
    const float *__restrict__ in = logits + ((long long)u * Tmax + t) * V;
    auto load = [&](int v) -> float {
        if (v >= V) return -INFINITY;
        float a = __ldg(in + v);
        return apply_relu ? fmaxf(a, 0.0f) : a;
    };

    if (kCached) {
        float a[4];
        const int v0 = lane * 4;
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = (v0 < Vp) ? load(v0 + k) : -INFINITY; m = fmaxf(m, a[k]); }
        m = warp_max(m);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) s += (a[k] == -INFINITY) ? 0.0f : expf(a[k] - m);
        s = warp_sum(s);
        const float ls = logf(s);
        if (lane < groups) {
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = (v0 + k < V) ? (a[k] - m) - ls : E2E_CTC_LOGZERO;
            out[lane] = make_float4(o[0], o[1], o[2], o[3]);
        }
    } else {
        float m = -INFINITY;
        for (int v = lane; v < V; v += 32) m = fmaxf(m, load(v));
        m = warp_max(m);
        float s = 0.0f;
        for (int v = lane; v < V; v += 32) s += expf(load(v) - m);
        s = warp_sum(s);
        const float ls = logf(s);
        for (int g = lane; g < groups; g += 32) {
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = g * 4 + k;
                o[k] = (v < V) ? (load(v) - m) - ls : E2E_CTC_LOGZERO;
            }
            out[g] = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// Large vocabularies (Vp > 128, e.g. the 10k subword vocabulary of BASELINE cfg3): one CTA per output row.  The row is
// read from HBM ONCE with 16-byte loads into shared memory (ReLU applied on the way), max and sum-exp are block
// reductions over the staged copy, and the log-probabilities leave as 16-byte stores: 8*V bytes of traffic per
// frame-row, the algorithmic minimum, instead of three scalar passes over the row.
constexpr int kWideThreads = 256;

__global__ void __launch_bounds__(kWideThreads)
ctc_log_softmax_wide_kernel(const float *__restrict__ logits, int U, int Tmax, int V, const int *__restrict__ enc_len,
                            int apply_relu, float *__restrict__ x, int Vp)
{
    extern __shared__ __align__(16) float s_row[];        // [Vp]
    __shared__ float s_red[kWideThreads / 32];
    __shared__ float s_bcast;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long row = blockIdx.x;                      // = t*U + u
    const int t = (int)(row / U), u = (int)(row % U);
    float4 *__restrict__ out = reinterpret_cast<float4 *>(x + row * Vp);
    const int groups = Vp >> 2;
    const int tu = enc_len ? enc_len[u] : Tmax;
    if (t >= tu) {
        const float4 z = make_float4(E2E_CTC_LOGZERO, E2E_CTC_LOGZERO, E2E_CTC_LOGZERO, E2E_CTC_LOGZERO);
        for (int g = tid; g < groups; g += kWideThreads) out[g] = z;
        return;
    }
    const float *__restrict__ in = logits + ((long long)u * Tmax + t) * V;
    const bool vec = ((V & 3) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
    float m = -INFINITY;
    if (vec) {
        const float4 *in4 = reinterpret_cast<const float4 *>(in);
        for (int g = tid; g < (V >> 2); g += kWideThreads) {
            float4 a = __ldg(in4 + g);
            if (apply_relu) { a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f); }
            reinterpret_cast<float4 *>(s_row)[g] = a;
            m = fmaxf(fmaxf(m, fmaxf(a.x, a.y)), fmaxf(a.z, a.w));
        }
    } else {
        for (int v = tid; v < V; v += kWideThreads) {
            float a = __ldg(in + v);
            if (apply_relu) a = fmaxf(a, 0.0f);
            s_row[v] = a;
            m = fmaxf(m, a);
        }
    }
    m = warp_max(m);
    if (lane == 0) s_red[wid] = m;
    __syncthreads();
    if (tid == 0) {
        float mm = s_red[0];
        for (int w = 1; w < kWideThreads / 32; ++w) mm = fmaxf(mm, s_red[w]);
        s_bcast = mm;
    }
    __syncthreads();
    m = s_bcast;
    float sum = 0.0f;
    for (int v = tid; v < V; v += kWideThreads) sum += expf(s_row[v] - m);
    sum = warp_sum(sum);
    __syncthreads();                                       // s_red / s_bcast are reused
    if (lane == 0) s_red[wid] = sum;
    __syncthreads();
    if (tid == 0) {
        float ss = 0.0f;
        for (int w = 0; w < kWideThreads / 32; ++w) ss += s_red[w];
        s_bcast = logf(ss);
    }
    __syncthreads();
    const float ls = s_bcast;
    for (int g = tid; g < groups; g += kWideThreads) {
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int v = g * 4 + k;
            o[k] = (v < V) ? (s_row[v] - m) - ls : E2E_CTC_LOGZERO;
        }
        out[g] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// One warp per utterance.  The reference's running sum is sequential fp32 (src/ctc.py:24-26), so the adds are kept in
// that order: the lanes load 32 frames' blank scores at once (independent loads), every lane then replays the 32 adds
// from shuffles (the chain is one FADD per frame) and keeps the partial sum of its own frame, and the 32 states leave
// as one 256-byte store.  r0: [U][Tmax][1][2].
constexpr int kInitWarps = 4;
__global__ void __launch_bounds__(kInitWarps * 32)
ctc_init_state_kernel(const float *__restrict__ x, int Tmax, int U, int Vp, const int *__restrict__ enc_len, float2 *__restrict__ r0)
{
    const int u = blockIdx.x * kInitWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (u >= U) return;
    const int tu = enc_len ? min(enc_len[u], Tmax) : Tmax;
    const float *xb = x + (long long)u * Vp + E2E_CTC_BLANK;
    float acc = 0.0f;
    float nxt = (lane < tu) ? __ldg(xb + (long long)lane * U * Vp) : 0.0f;
    for (int t0 = 0; t0 < tu; t0 += 32) {
        const float cur = nxt;
        if (t0 + 32 + lane < tu) nxt = __ldg(xb + (long long)(t0 + 32 + lane) * U * Vp);      // next group in flight under the chain
        float mine = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float b = __shfl_sync(0xffffffffu, cur, i);
            acc = (t0 + i == 0) ? b : __fadd_rn(acc, b);
            if (i == lane) mine = acc;
        }
        if (t0 + lane < tu) r0[(long long)u * Tmax + t0 + lane] = make_float2(E2E_CTC_LOGZERO, mine);
    }
}

}  // namespace e2e

extern "C" int e2e_ctc_log_softmax(const float *logits, int U, int Tmax, int V, const int *enc_len,
                                   int apply_relu, float *x, int Vp, void *stream)
{
    using namespace e2e;
    if (!logits || !x || U <= 0 || Tmax <= 0 || V <= 0) return set_error(E2E_ERR_ARG, "e2e_ctc_log_softmax: bad argument");
    if (Vp != e2e_padded_vocab(V)) return set_error(E2E_ERR_ARG, "e2e_ctc_log_softmax: Vp=%d, expected %d", Vp, e2e_padded_vocab(V));
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return set_error(E2E_ERR_ARG, "e2e_ctc_log_softmax: x must be 16-byte aligned");
    const long long rows = (long long)Tmax * U;
    const unsigned blocks = (unsigned)((rows + kRowWarps - 1) / kRowWarps);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (Vp <= 128) {
        const int groups = Vp / 4;
        const int G = groups <= 1 ? 1 : groups <= 2 ? 2 : groups <= 4 ? 4 : groups <= 8 ? 8 : groups <= 16 ? 16 : 32;
        const long long per_cta = (long long)8 * (32 / G);                      // 8 warps per CTA
        const unsigned nb = (unsigned)((rows + per_cta - 1) / per_cta);
        switch (G) {
            case 1: ctc_log_softmax_narrow_kernel<1><<<nb, 256, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp); break;
            case 2: ctc_log_softmax_narrow_kernel<2><<<nb, 256, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp); break;
            case 4: ctc_log_softmax_narrow_kernel<4><<<nb, 256, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp); break;
            case 8: ctc_log_softmax_narrow_kernel<8><<<nb, 256, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp); break;
            case 16: ctc_log_softmax_narrow_kernel<16><<<nb, 256, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp); break;
            default: ctc_log_softmax_narrow_kernel<32><<<nb, 256, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp); break;
        }
    } else if ((size_t)Vp * 4 <= 160 * 1024 && rows <= 0x7fffffffLL) {
        const size_t smem = (size_t)Vp * 4;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(ctc_log_softmax_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }
        ctc_log_softmax_wide_kernel<<<(unsigned)rows, kWideThreads, smem, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp);
    } else
        ctc_log_softmax_kernel<false><<<blocks, kRowWarps * 32, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp);
    count_launch();
    return check_launch("e2e_ctc_log_softmax");
}

extern "C" int e2e_ctc_init_state(const float *x, int Tmax, int U, int Vp, const int *enc_len, float *r0, void *stream)
{
    using namespace e2e;
    if (!x || !r0 || U <= 0 || Tmax <= 0 || Vp <= 0) return set_error(E2E_ERR_ARG, "e2e_ctc_init_state: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ctc_init_state_kernel<<<(U + kInitWarps - 1) / kInitWarps, kInitWarps * 32, 0, st>>>(x, Tmax, U, Vp, enc_len, reinterpret_cast<float2 *>(r0));
    count_launch();
    return check_launch("e2e_ctc_init_state");
}
