// Fused location-aware attention energy + masked softmax.
// Replaces, for one decode step and every live hypothesis n = u*B + b,
//     loc    = tanh(loc_proj(conv_feat^T))                       src/module.py:1163
//     energy = gen_energy(tanh(key + query + loc))               src/module.py:1168
//     attn   = softmax(mask(energy / temperature))               src/module.py:1109-1113
// i.e. everything that would otherwise materialise [U*B, T, A] temporaries several times per
// step.  The convolution over the previous alignment (cuDNN) and the context product
// attn x value (cuBLAS batched GEMM) stay library calls.
//
// Mapping: one CTA per hypothesis, one THREAD per encoder frame t (256 frames per pass).  The
// per-channel constants (loc_proj row, query[n][a], gen_energy weight) are identical for every
// lane, so they sit in shared memory as 4 x float4 per channel and are read as broadcasts; the
// key row of a frame is streamed with 16-byte loads; nothing is reduced across lanes until the
// softmax.  tanh(x) = 1 - 2/(1 + exp(2x)) through MUFU.EX2 + MUFU.RCP (absolute error ~2e-7).
#include "common.cuh"

namespace e2e {

constexpr int kAttThreads = 256;
constexpr int kAttMaxK = 12;     // loc_kernel_num <= 12 (10 in every shipped config)

__device__ __forceinline__ float tanh_mufu(float x)
{
    const float t = ex2_approx(x * 2.8853900817779268f);      // exp(2x)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
    return fmaf(-2.0f, r, 1.0f);
}

struct AttParams {
    const float *key, *query, *loc_feat; const int *enc_len;
    const float *w_proj, *w_energy; float b_energy, temperature;
    int B, T, A, K; float *attn;
};

__global__ void __launch_bounds__(kAttThreads)
attention_loc_kernel(const AttParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *cst = reinterpret_cast<float4 *>(smem_raw);                  // [A][4]: w_proj[a][0..11] | q[a], w_e[a], 0, 0
    float *e_s = reinterpret_cast<float *>(smem_raw + (size_t)p.A * 64);  // [T]
    float *red = e_s + ((p.T + 3) & ~3);                                  // [32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x, u = n / p.B;
    const int T = p.T, A = p.A, K = p.K;
    const int Tu = min(p.enc_len[u], T);

    for (int a = tid; a < A; a += kAttThreads) {
        float w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) w[k] = 0.0f;
#pragma unroll
        for (int k = 0; k < kAttMaxK; ++k)
            if (k < K) w[k] = __ldg(p.w_proj + (size_t)a * K + k);
        w[12] = __ldg(p.query + (size_t)n * A + a);
        w[13] = __ldg(p.w_energy + a);
#pragma unroll
        for (int q = 0; q < 4; ++q) cst[a * 4 + q] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
    __syncthreads();

    for (int t = tid; t < T; t += kAttThreads) {
        float s = -INFINITY;
        if (t < Tu) {
            float f[kAttMaxK];
#pragma unroll
            for (int k = 0; k < kAttMaxK; ++k) f[k] = (k < K) ? __ldg(p.loc_feat + ((size_t)n * K + k) * T + t) : 0.0f;
            const float4 *krow = reinterpret_cast<const float4 *>(p.key + ((size_t)u * T + t) * A);
            float acc = p.b_energy;
            for (int a4 = 0; a4 < A / 4; ++a4) {
                const float4 kv = __ldg(krow + a4);
                const float kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 c0 = cst[(a4 * 4 + i) * 4 + 0], c1 = cst[(a4 * 4 + i) * 4 + 1];
                    const float4 c2 = cst[(a4 * 4 + i) * 4 + 2], c3 = cst[(a4 * 4 + i) * 4 + 3];
                    float loc = c0.x * f[0];
                    loc = fmaf(c0.y, f[1], loc); loc = fmaf(c0.z, f[2], loc); loc = fmaf(c0.w, f[3], loc);
                    loc = fmaf(c1.x, f[4], loc); loc = fmaf(c1.y, f[5], loc); loc = fmaf(c1.z, f[6], loc); loc = fmaf(c1.w, f[7], loc);
                    loc = fmaf(c2.x, f[8], loc); loc = fmaf(c2.y, f[9], loc); loc = fmaf(c2.z, f[10], loc); loc = fmaf(c2.w, f[11], loc);
                    const float x = (kk[i] + c3.x) + tanh_mufu(loc);
                    acc = fmaf(c3.y, tanh_mufu(x), acc);
                }
            }
            s = __fdiv_rn(acc, p.temperature);
        }
        e_s[t] = s;
    }
    __syncthreads();

    // masked softmax over t (module.py:1109-1113); frames t >= enc_len get exactly 0
    float m = -INFINITY;
    for (int t = tid; t < T; t += kAttThreads) m = fmaxf(m, e_s[t]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < kAttThreads / 32; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float sum = 0.0f;
    for (int t = tid; t < T; t += kAttThreads) {
        const float e = (t < Tu) ? expf(e_s[t] - m) : 0.0f;
        e_s[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.0f;
    for (int w = 0; w < kAttThreads / 32; ++w) sum += red[w];
    float *out = p.attn + (size_t)n * T;
    for (int t = tid; t < T; t += kAttThreads) out[t] = __fdiv_rn(e_s[t], sum);
}

}  // namespace e2e

extern "C" int e2e_attention_loc_step(const float *key, const float *query, const float *loc_feat, const int *enc_len,
                                      const float *w_proj, const float *w_energy, float b_energy, float temperature,
                                      int n_hyp, int B, int T, int A, int K, float *attn, void *stream)
{
    using namespace e2e;
    if (!key || !query || !loc_feat || !enc_len || !w_proj || !w_energy || !attn)
        return set_error(E2E_ERR_ARG, "e2e_attention_loc_step: null pointer");
    if (n_hyp <= 0 || B <= 0 || T <= 0 || A <= 0 || K <= 0) return set_error(E2E_ERR_ARG, "e2e_attention_loc_step: bad size");
    if (K > kAttMaxK || (A & 3) != 0) return set_error(E2E_ERR_UNSUPPORTED, "e2e_attention_loc_step: needs loc_kernel_num <= 12 and dim %% 4 == 0");
    if (reinterpret_cast<uintptr_t>(key) & 15) return set_error(E2E_ERR_ARG, "e2e_attention_loc_step: key must be 16-byte aligned");
    const size_t smem = (size_t)A * 64 + (size_t)((T + 3) & ~3) * 4 + 32 * 4;
    if (smem > 200 * 1024) return set_error(E2E_ERR_UNSUPPORTED, "e2e_attention_loc_step: %zu bytes of shared memory needed", smem);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(attention_loc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    AttParams p;
    p.key = key; p.query = query; p.loc_feat = loc_feat; p.enc_len = enc_len; p.w_proj = w_proj; p.w_energy = w_energy;
    p.b_energy = b_energy; p.temperature = temperature; p.B = B; p.T = T; p.A = A; p.K = K; p.attn = attn;
    attention_loc_kernel<<<n_hyp, kAttThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    count_launch();
    return check_launch("e2e_attention_loc_step");
}
