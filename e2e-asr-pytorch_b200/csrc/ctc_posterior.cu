// Kernel (1): fused ReLU + log-softmax of the CTC head, written frame-major [Tmax][U][Vp]
// with 16-byte vector stores, and the blank running sum that seeds the prefix states.
// Replaces src/decode.py:94-95 (minus the cuBLAS Linear) and src/ctc.py:19-27.
#include "common.cuh"

namespace e2e {

constexpr int kRowWarps = 8;   // rows (one per warp) per CTA

// One warp per output row (t,u).  Lane l owns the 4-column groups l, l+32, ... of the row.
// kCached: the whole row fits in one group per lane (Vp <= 128) and stays in registers.
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

// One thread per utterance: the reference's running sum is sequential fp32 (src/ctc.py:24-26),
// so the adds are kept in that order.  r0: [U][Tmax][1][2].
__global__ void ctc_init_state_kernel(const float *__restrict__ x, int Tmax, int U, int Vp,
                                      const int *__restrict__ enc_len, float2 *__restrict__ r0)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    const int tu = enc_len ? enc_len[u] : Tmax;
    float acc = 0.0f;
    for (int t = 0; t < tu; ++t) {
        const float b = __ldg(x + ((long long)t * U + u) * Vp + E2E_CTC_BLANK);
        acc = (t == 0) ? b : __fadd_rn(acc, b);
        r0[(long long)u * Tmax + t] = make_float2(E2E_CTC_LOGZERO, acc);
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
    if (Vp <= 128)
        ctc_log_softmax_kernel<true><<<blocks, kRowWarps * 32, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp);
    else
        ctc_log_softmax_kernel<false><<<blocks, kRowWarps * 32, 0, st>>>(logits, U, Tmax, V, enc_len, apply_relu, x, Vp);
    count_launch();
    return check_launch("e2e_ctc_log_softmax");
}

extern "C" int e2e_ctc_init_state(const float *x, int Tmax, int U, int Vp, const int *enc_len, float *r0, void *stream)
{
    using namespace e2e;
    if (!x || !r0 || U <= 0 || Tmax <= 0 || Vp <= 0) return set_error(E2E_ERR_ARG, "e2e_ctc_init_state: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ctc_init_state_kernel<<<(U + 127) / 128, 128, 0, st>>>(x, Tmax, U, Vp, enc_len, reinterpret_cast<float2 *>(r0));
    count_launch();
    return check_launch("e2e_ctc_init_state");
}
