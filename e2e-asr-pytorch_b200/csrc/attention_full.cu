// Whole location-aware attention step in one kernel:
//     feat   = loc_conv(prev_att)                                 src/module.py:1163 (Conv1d 1->K, 2P+1 taps, zero padded)
//     loc    = tanh(loc_proj(feat^T))                             src/module.py:1163
//     energy = gen_energy(tanh(key + query + loc))                src/module.py:1168
//     attn   = softmax(mask(energy / temperature))                src/module.py:1109-1113
//     ctx    = attn x value                                       src/module.py:1114
// for every live hypothesis n = u*B + b of the first n_run utterances.  Nothing of shape
// [U*B, T, *] is ever materialised: per step the kernel reads key/value of each utterance,
// the previous alignments and the queries, and writes the new alignments and contexts.
//
// Mapping: one CTA per (utterance, group of NB beam slots), one THREAD per encoder frame t
// (256 frames per pass) that carries the NB hypotheses of the group together, so a key value is
// fetched once per group (the keys arrive channel-major, key_t[u][a][t], so the warp's 32 frames
// are one coalesced 128-byte load per channel) and the per-channel constants (loc_proj row, gen_energy weight, the
// NB queries) are shared-memory broadcasts.  Frames t >= enc_len[u] are never touched: a batch
// padded to the longest utterance costs nothing.  tanh(x) = 1 - 2/(1 + exp(2x)) through
// MUFU.EX2 + MUFU.RCP (absolute error ~2e-7); the kernel is MUFU-bound by construction
// (4 MUFU per (hypothesis, frame, channel)).
// The context product runs in the same CTA afterwards: thread <-> 4 output columns (16-byte
// loads, 4 frames in flight), the NB alignments are read as one broadcast LDS.128 per frame.
// Results do not depend on NB (every hypothesis sees the same operations in the same order),
// so the host picks NB by launch size only.
#include "common.cuh"

namespace e2e {

constexpr int kAfThreads = 256;
constexpr int kAfMaxK = 12;

__device__ __forceinline__ float af_tanh(float x)
{
    const float t = ex2_approx(x * 2.8853900817779268f);      // exp(2x)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    return fmaf(-2.0f, r, 1.0f);
}

struct AttFullParams {
    const float *key_t, *value, *query, *prev_att; const int *enc_len;
    const float *w_conv, *w_proj, *w_energy; float b_energy, temperature;
    int B, T, A, K, W, E, groups;
    float *attn, *ctx;
};

// shared memory (floats): pa [NB][TP] | wc [W][KP] | cst [A][20] | es [T4][NB] | red [8*NB]
template <int NB, int KP>
__global__ void __launch_bounds__(kAfThreads)
attention_full_kernel(const AttFullParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int u = blockIdx.x / p.groups, g = blockIdx.x - u * p.groups;
    const int T = p.T, A = p.A, K = p.K, W = p.W, P = W / 2;
    const int Tu = min(p.enc_len[u], T);
    const int b0 = g * NB;                                    // first beam slot of this group
    const int TP = (T + W - 1 + 3) & ~3;
    const int T4 = (T + 3) & ~3;
    float *pa = reinterpret_cast<float *>(smem_raw);
    float *wc = pa + NB * TP;
    float4 *cst = reinterpret_cast<float4 *>(wc + ((W * KP + 3) & ~3));
    float *es = reinterpret_cast<float *>(cst + (size_t)A * 5);
    float *red = es + (size_t)T4 * NB;

    // ---- stage: previous alignments with a zero halo, filters, per-channel constants ----------
    for (int i = tid; i < NB * TP; i += kAfThreads) {
        const int b = i / TP, tt = i - b * TP - P;
        float v = 0.0f;
        if (b0 + b < p.B && tt >= 0 && tt < Tu) v = __ldg(p.prev_att + ((size_t)u * p.B + b0 + b) * T + tt);
        pa[i] = v;
    }
    for (int i = tid; i < W * KP; i += kAfThreads) {
        const int j = i / KP, k = i - j * KP;
        wc[i] = (k < K) ? __ldg(p.w_conv + (size_t)k * W + j) : 0.0f;
    }
    for (int a = tid; a < A; a += kAfThreads) {
        float w[kAfMaxK];
#pragma unroll
        for (int k = 0; k < kAfMaxK; ++k) w[k] = (k < K) ? __ldg(p.w_proj + (size_t)a * K + k) : 0.0f;
        float q[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int b = 0; b < NB; ++b)
            if (b0 + b < p.B) q[b] = __ldg(p.query + ((size_t)u * p.B + b0 + b) * A + a);
        cst[a * 5 + 0] = make_float4(w[0], w[1], w[2], w[3]);
        cst[a * 5 + 1] = make_float4(w[4], w[5], w[6], w[7]);
        cst[a * 5 + 2] = make_float4(w[8], w[9], w[10], w[11]);
        cst[a * 5 + 3] = make_float4(q[0], q[1], q[2], q[3]);
        cst[a * 5 + 4] = make_float4(__ldg(p.w_energy + a), 0.0f, 0.0f, 0.0f);
    }
    __syncthreads();

    // ---- energies ---------------------------------------------------------------------------------
    for (int t = tid; t < T; t += kAfThreads) {
        float s[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) s[b] = -INFINITY;
        if (t < Tu) {
            // location features: f[b][k] = sum_j w_conv[k][j] * prev_att[b][t + j - P]
            float f[NB][KP];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int k = 0; k < KP; ++k) f[b][k] = 0.0f;
            const float *pat = pa + t;
#pragma unroll 2
            for (int j = 0; j < W; ++j) {
                float wj[KP];
#pragma unroll
                for (int k2 = 0; k2 < KP / 2; ++k2) {
                    const float2 w2 = *reinterpret_cast<const float2 *>(wc + j * KP + 2 * k2);
                    wj[2 * k2] = w2.x; wj[2 * k2 + 1] = w2.y;
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float a = pat[b * TP + j];
#pragma unroll
                    for (int k = 0; k < KP; ++k) f[b][k] = fmaf(wj[k], a, f[b][k]);
                }
            }
            const float *kcol = p.key_t + (size_t)u * A * T + t;        // key_t[u][a][t]: coalesced across the warp's frames
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b] = p.b_energy;
            for (int a4 = 0; a4 < A / 4; ++a4) {
                float kk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) kk[i] = __ldg(kcol + (size_t)(a4 * 4 + i) * T);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 *c = cst + (a4 * 4 + i) * 5;
                    const float4 c0 = c[0], c1 = c[1], c2 = c[2], cq = c[3];
                    const float we = c[4].x;
                    const float wp[kAfMaxK] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w};
                    const float qq[4] = {cq.x, cq.y, cq.z, cq.w};
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        float loc = wp[0] * f[b][0];
#pragma unroll
                        for (int k = 1; k < KP; ++k) loc = fmaf(wp[k], f[b][k], loc);
                        const float x = (kk[i] + qq[b]) + af_tanh(loc);
                        acc[b] = fmaf(we, af_tanh(x), acc[b]);
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < NB; ++b) s[b] = __fdiv_rn(acc[b], p.temperature);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) es[t * NB + b] = s[b];
    }
    __syncthreads();

    // ---- masked softmax over t, one hypothesis after the other (module.py:1109-1113) ---------------
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float m = -INFINITY;
        for (int t = tid; t < T; t += kAfThreads) m = fmaxf(m, es[t * NB + b]);
        m = warp_max(m);
        if (lane == 0) red[warp] = m;
        __syncthreads();
        m = red[0];
        for (int w = 1; w < kAfThreads / 32; ++w) m = fmaxf(m, red[w]);
        __syncthreads();
        float sum = 0.0f;
        for (int t = tid; t < T; t += kAfThreads) {
            const float e = (t < Tu) ? expf(es[t * NB + b] - m) : 0.0f;
            es[t * NB + b] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        sum = 0.0f;
        for (int w = 0; w < kAfThreads / 32; ++w) sum += red[w];
        const bool real = b0 + b < p.B;
        float *out = p.attn + ((size_t)u * p.B + b0 + b) * T;
        for (int t = tid; t < T; t += kAfThreads) {
            const float a = __fdiv_rn(es[t * NB + b], sum);
            es[t * NB + b] = a;
            if (real) out[t] = a;
        }
        __syncthreads();
    }

    // ---- context: ctx[b][e] = sum_t attn[b][t] * value[u][t][e]  (module.py:1114) -------------------
    // thread <-> 4 consecutive columns (one 16-byte load per frame, whole row coalesced), 4 frames in flight
    const float *vbase = p.value + (size_t)u * T * p.E;
    if ((p.E & 3) == 0) {
        for (int e = tid * 4; e < p.E; e += kAfThreads * 4) {
            float acc[NB][4];
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[b][i] = 0.0f;
            const float *vp = vbase + e;
            int t = 0;
            for (; t + 4 <= Tu; t += 4) {
                float4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float4 *>(vp + (size_t)(t + q) * p.E));
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const float a = es[(t + q) * NB + b];
                        acc[b][0] = fmaf(a, v[q].x, acc[b][0]); acc[b][1] = fmaf(a, v[q].y, acc[b][1]);
                        acc[b][2] = fmaf(a, v[q].z, acc[b][2]); acc[b][3] = fmaf(a, v[q].w, acc[b][3]);
                    }
            }
            for (; t < Tu; ++t) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(vp + (size_t)t * p.E));
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const float a = es[t * NB + b];
                    acc[b][0] = fmaf(a, v.x, acc[b][0]); acc[b][1] = fmaf(a, v.y, acc[b][1]);
                    acc[b][2] = fmaf(a, v.z, acc[b][2]); acc[b][3] = fmaf(a, v.w, acc[b][3]);
                }
            }
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (b0 + b < p.B)
                    *reinterpret_cast<float4 *>(p.ctx + ((size_t)u * p.B + b0 + b) * p.E + e) = make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]);
        }
    } else {
        for (int e = tid; e < p.E; e += kAfThreads) {
            float acc[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) acc[b] = 0.0f;
            const float *vp = vbase + e;
            for (int t = 0; t < Tu; ++t) {
                const float v = __ldg(vp + (size_t)t * p.E);
#pragma unroll
                for (int b = 0; b < NB; ++b) acc[b] = fmaf(es[t * NB + b], v, acc[b]);
            }
#pragma unroll
            for (int b = 0; b < NB; ++b)
                if (b0 + b < p.B) p.ctx[((size_t)u * p.B + b0 + b) * p.E + e] = acc[b];
        }
    }
}

static size_t att_full_smem(int NB, int KP, int T, int A, int W)
{
    const size_t TP = (size_t)((T + W - 1 + 3) & ~3), T4 = (size_t)((T + 3) & ~3);
    return ((size_t)NB * TP + (size_t)((W * KP + 3) & ~3) + (size_t)A * 20 + T4 * NB + 8 * NB + 8) * 4;
}

template <int NB>
static int att_full_launch(const AttFullParams &p, int KP, int n_run, cudaStream_t st)
{
    void (*kern)(AttFullParams) = nullptr;
    switch (KP) {
        case 4: kern = attention_full_kernel<NB, 4>; break;
        case 8: kern = attention_full_kernel<NB, 8>; break;
        case 10: kern = attention_full_kernel<NB, 10>; break;
        default: kern = attention_full_kernel<NB, 12>; break;
    }
    const size_t smem = att_full_smem(NB, KP, p.T, p.A, p.W);
    if (smem > 200 * 1024) return set_error(E2E_ERR_UNSUPPORTED, "e2e_attention_loc_full: %zu bytes of shared memory needed", smem);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<(unsigned)(n_run * p.groups), kAfThreads, smem, st>>>(p);
    count_launch();
    return check_launch("e2e_attention_loc_full");
}

}  // namespace e2e

extern "C" int e2e_attention_loc_full(const float *key_t, const float *value, const float *query, const float *prev_att,
                                      const int *enc_len, const float *w_conv, const float *w_proj, const float *w_energy,
                                      float b_energy, float temperature, int n_run, int B, int T, int A, int K, int W, int E,
                                      int hyps_per_cta, float *attn, float *ctx, void *stream)
{
    using namespace e2e;
    if (!key_t || !value || !query || !prev_att || !enc_len || !w_conv || !w_proj || !w_energy || !attn || !ctx)
        return set_error(E2E_ERR_ARG, "e2e_attention_loc_full: null pointer");
    if (n_run <= 0 || B <= 0 || T <= 0 || A <= 0 || K <= 0 || W <= 0 || E <= 0) return set_error(E2E_ERR_ARG, "e2e_attention_loc_full: bad size");
    if (K > kAfMaxK || (A & 3) != 0 || (W & 1) == 0)
        return set_error(E2E_ERR_UNSUPPORTED, "e2e_attention_loc_full: needs loc_kernel_num <= 12, dim %% 4 == 0 and an odd filter length");
    if ((reinterpret_cast<uintptr_t>(value) & 15) || (reinterpret_cast<uintptr_t>(ctx) & 15))
        return set_error(E2E_ERR_ARG, "e2e_attention_loc_full: value and ctx must be 16-byte aligned");
    int NB = hyps_per_cta;
    if (NB <= 0) {   // by launch size: enough CTAs to cover the machine, otherwise as much key reuse as possible
        const long long want = 2LL * 148;
        NB = ((long long)n_run * ((B + 3) / 4) >= want || B == 1) ? 4 : (((long long)n_run * ((B + 1) / 2) >= want) ? 2 : 1);
    }
    if (NB != 1 && NB != 2 && NB != 4) return set_error(E2E_ERR_ARG, "e2e_attention_loc_full: hyps_per_cta must be 0, 1, 2 or 4");
    while (NB > 1 && NB / 2 >= B) NB /= 2;
    AttFullParams p;
    p.key_t = key_t; p.value = value; p.query = query; p.prev_att = prev_att; p.enc_len = enc_len;
    p.w_conv = w_conv; p.w_proj = w_proj; p.w_energy = w_energy; p.b_energy = b_energy; p.temperature = temperature;
    p.B = B; p.T = T; p.A = A; p.K = K; p.W = W; p.E = E; p.groups = (B + NB - 1) / NB;
    p.attn = attn; p.ctx = ctx;
    const int KP = K <= 4 ? 4 : (K <= 8 ? 8 : (K <= 10 ? 10 : 12));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (NB == 4) return att_full_launch<4>(p, KP, n_run, st);
    if (NB == 2) return att_full_launch<2>(p, KP, n_run, st);
    return att_full_launch<1>(p, KP, n_run, st);
}
