// 3x3 "same" convolution of the VGG front end as an fp32-accurate tensor-core GEMM
// (the convolutions are src/module.py:672-686).
// The reference runs these layers as cuDNN fp32 SIMT convolutions: 62 TFLOP per pass over the
// bench workload, a quarter of the whole decode.  Here the activations stay NHWC and a block of
// pixels is unfolded into the A operand of the library's bf16 GEMM as the exact 3-piece bf16 split
// of every fp32 value ([a1 | a2 | a3], see lstm_step.cu / stepper.SplitLinear), so that
//     y[p][co] = sum_k A[p][k] * Wmat[k][co],   k = (dy*3 + dx)*C + c,  Wmat = weight.permute(2,3,1,0)
// runs on the tensor cores with fp32 accumulation and loses nothing against the fp32 SIMT result.
// This kernel is the unfold + split: one pass over the (L2-resident) input block, 16-byte loads,
// 8-byte stores.  Rows h >= valid[n] of an utterance read as zero, which is exactly the masking
// VGGFrontEnd.forward_masked applies between layers (a padded batch row must see the zero padding a
// batch-1 call would).
#include "common.cuh"
#include <cuda_bf16.h>

namespace e2e {

__device__ __forceinline__ void conv_split3(float x, __nv_bfloat16 &a1, __nv_bfloat16 &a2, __nv_bfloat16 &a3)
{
    a1 = __float2bfloat16_rn(x);
    const float r1 = __fsub_rn(x, __bfloat162float(a1));
    a2 = __float2bfloat16_rn(r1);
    a3 = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(a2)));
}

struct alignas(8) ConvBf16x4 { __nv_bfloat16 v[4]; };

// in  [N][H][W][C] fp32 (NHWC), valid [N] rows per image; pixels p0 .. p0+P-1 of the flattened (n,h,w) index
// out [P][3*9C] bf16
__global__ void __launch_bounds__(256)
im2col3x3_split_kernel(const float *__restrict__ in, const int *__restrict__ valid, int H, int W, int C,
                       long long p0, int P, __nv_bfloat16 *__restrict__ out)
{
    // one warp per output pixel (its coordinates are decoded once, warp uniform); the lanes sweep the
    // pixel's 9 taps x C/4 channel groups: 512-byte contiguous reads, 256-byte contiguous writes per piece
    const int C4 = C >> 2;
    const long long K = 9LL * C;
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    for (int pl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pl < P; pl += warps_per_grid) {
        const long long p = p0 + pl;
        const int w = (int)(p % W);
        const long long nh = p / W;
        const int h = (int)(nh % H);
        const int n = (int)(nh / H);
        const int vrows = __ldg(valid + n);
        const float *img = in + (long long)n * H * W * C;
        __nv_bfloat16 *orow = out + (long long)pl * 3 * K;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int hs = h + dy - 1;
            const bool row_ok = hs >= 0 && hs < H && hs < vrows;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ws = w + dx - 1;
                const bool ok = row_ok && ws >= 0 && ws < W;
                const float4 *src = reinterpret_cast<const float4 *>(img + ((long long)hs * W + ws) * C);
                __nv_bfloat16 *o = orow + (dy * 3 + dx) * C;
                for (int c4 = lane; c4 < C4; c4 += 32) {
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (ok) v = __ldg(src + c4);
                    ConvBf16x4 q0, q1, q2;
                    conv_split3(v.x, q0.v[0], q1.v[0], q2.v[0]);
                    conv_split3(v.y, q0.v[1], q1.v[1], q2.v[1]);
                    conv_split3(v.z, q0.v[2], q1.v[2], q2.v[2]);
                    conv_split3(v.w, q0.v[3], q1.v[3], q2.v[3]);
                    *reinterpret_cast<ConvBf16x4 *>(o + c4 * 4) = q0;
                    *reinterpret_cast<ConvBf16x4 *>(o + K + c4 * 4) = q1;
                    *reinterpret_cast<ConvBf16x4 *>(o + 2 * K + c4 * 4) = q2;
                }
            }
        }
    }
}

// y[p][c] = relu(y[p][c] + bias[c]) for valid rows, 0 for rows h >= valid[n]   (in place, NHWC)
__global__ void __launch_bounds__(256)
bias_relu_mask_kernel(float *__restrict__ y, const float *__restrict__ bias, const int *__restrict__ valid,
                      int H, int W, int C, long long p0, long long P)
{
    const int C4 = C >> 2;
    const long long total = P * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pl = i / C4;
        const int c4 = (int)(i - pl * C4);
        const long long nh = (p0 + pl) / W;
        const int h = (int)(nh % H);
        const int n = (int)(nh / H);
        float4 *ptr = reinterpret_cast<float4 *>(y + pl * C) + c4;
        float4 v = *ptr;
        if (h < __ldg(valid + n)) {
            const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
            v.x = fmaxf(__fadd_rn(v.x, b.x), 0.0f); v.y = fmaxf(__fadd_rn(v.y, b.y), 0.0f);
            v.z = fmaxf(__fadd_rn(v.z, b.z), 0.0f); v.w = fmaxf(__fadd_rn(v.w, b.w), 0.0f);
        } else {
            v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        *ptr = v;
    }
}

}  // namespace e2e

extern "C" int e2e_conv3x3_unfold_split(const float *in_nhwc, const int *valid_rows, int N, int H, int W, int C,
                                        long long first_pixel, int n_pixels, void *out_bf16, void *stream)
{
    using namespace e2e;
    if (!in_nhwc || !valid_rows || !out_bf16) return set_error(E2E_ERR_ARG, "e2e_conv3x3_unfold_split: null pointer");
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || n_pixels <= 0 || first_pixel < 0 ||
        first_pixel + n_pixels > (long long)N * H * W)
        return set_error(E2E_ERR_ARG, "e2e_conv3x3_unfold_split: bad size (C must be a multiple of 4)");
    if ((reinterpret_cast<uintptr_t>(in_nhwc) & 15) || (reinterpret_cast<uintptr_t>(out_bf16) & 7))
        return set_error(E2E_ERR_ARG, "e2e_conv3x3_unfold_split: misaligned buffer");
    long long blocks = ((long long)n_pixels + 7) / 8;           // 8 warps per block, one pixel per warp at a time
    if (blocks > 148LL * 64) blocks = 148LL * 64;             // grid-stride: a few waves of the machine
    im2col3x3_split_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in_nhwc, valid_rows, H, W, C, first_pixel, n_pixels, static_cast<__nv_bfloat16 *>(out_bf16));
    count_launch();
    return check_launch("e2e_conv3x3_unfold_split");
}

extern "C" int e2e_conv_bias_relu_mask(float *y_nhwc, const float *bias, const int *valid_rows, int N, int H, int W, int C,
                                       long long first_pixel, long long n_pixels, void *stream)
{
    using namespace e2e;
    if (!y_nhwc || !bias || !valid_rows) return set_error(E2E_ERR_ARG, "e2e_conv_bias_relu_mask: null pointer");
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || n_pixels <= 0 || first_pixel < 0 ||
        first_pixel + n_pixels > (long long)N * H * W)
        return set_error(E2E_ERR_ARG, "e2e_conv_bias_relu_mask: bad size (C must be a multiple of 4)");
    if ((reinterpret_cast<uintptr_t>(y_nhwc) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15))
        return set_error(E2E_ERR_ARG, "e2e_conv_bias_relu_mask: misaligned buffer");
    const long long total = n_pixels * (C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    bias_relu_mask_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        y_nhwc, bias, valid_rows, H, W, C, first_pixel, n_pixels);
    count_launch();
    return check_launch("e2e_conv_bias_relu_mask");
}

// ---------------------------------------------------------------------------------------------
// First VGG layer (Cin -> Cout, 3x3 "same", bias, ReLU; src/module.py:672-674) straight from the
// feature frames into NHWC, and bias + ReLU + mask fused with the 2x2 ceil-mode max pooling
// (src/module.py:676,683) for the layers that are followed by one.
// ---------------------------------------------------------------------------------------------
namespace e2e {

// feat [N][L][Cin*F] (frame t = Cin blocks of F frequency bins, src/module.py:688-690), weight [Cout][Cin][3][3],
// out [N][L][F][Cout] NHWC = relu(conv + bias).  Input rows t >= valid[n] read as zero.  One CTA per (n, t) row:
// the three input rows sit in shared memory, the weights as [tap][cin][cout] (one LDS.128 per 4 output channels);
// lane <-> 4 output channels, so a warp writes one pixel's channels as 512 contiguous bytes.
constexpr int kC1Rows = 4;       // output rows per CTA (the weights are staged once per CTA)

__global__ void __launch_bounds__(256)
conv1_direct_kernel(const float *__restrict__ feat, long long feat_pitch_n, const float *__restrict__ weight,
                    const float *__restrict__ bias, const int *__restrict__ valid, int L, int F, int Cin, int Cout,
                    float *__restrict__ out)
{
    extern __shared__ __align__(16) float c1_smem[];
    float *ws = c1_smem;                                   // [9*Cin][Cout]
    float *rows = ws + 9 * Cin * Cout;                     // [kC1Rows + 2][Cin][F + 5] (zero halo in frequency + pad)
    const int n = blockIdx.y, t0 = blockIdx.x * kC1Rows, tid = threadIdx.x;
    const int Fp = F + 5;
    for (int i = tid; i < 9 * Cin * Cout; i += blockDim.x) {
        const int co = i % Cout, r = i / Cout;             // r = tap*Cin + ci
        const int ci = r % Cin, tap = r / Cin;
        ws[i] = __ldg(weight + ((size_t)co * Cin + ci) * 9 + tap);
    }
    const int vr = __ldg(valid + n);
    for (int i = tid; i < (kC1Rows + 2) * Cin * Fp; i += blockDim.x) {
        const int f = i % Fp - 1, r = i / Fp;              // r = row*Cin + ci
        const int ci = r % Cin, row = r / Cin;
        const int ts = t0 + row - 1;
        float v = 0.0f;
        if (f >= 0 && f < F && ts >= 0 && ts < L && ts < vr) v = __ldg(feat + (size_t)n * feat_pitch_n + (size_t)ts * Cin * F + ci * F + f);
        rows[i] = v;
    }
    __syncthreads();
    // lane <-> 4 output channels, warp <-> a run of 4 pixels of one row: per (dy, cin) 6 broadcast inputs and
    // 3 weight quads feed 48 FMAs
    const int C4 = Cout >> 2, FG = (F + 3) >> 2;
    const int lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    for (int c4 = lane; c4 < C4; c4 += 32) {
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
        for (int item = warp; item < kC1Rows * FG; item += n_warps) {
            const int tr = item / FG, f0 = (item - tr * FG) * 4;
            const int t = t0 + tr;
            if (t >= L) break;
            float4 acc[4] = {b4, b4, b4, b4};
            for (int dy = 0; dy < 3; ++dy)
                for (int ci = 0; ci < Cin; ++ci) {
                    const float *rp = rows + ((tr + dy) * Cin + ci) * Fp + f0;       // bins f0-1 .. f0+4 with the halo offset
                    float x[6];
#pragma unroll
                    for (int q = 0; q < 6; ++q) x[q] = rp[q];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const float4 w = *reinterpret_cast<const float4 *>(ws + ((dy * 3 + dx) * Cin + ci) * Cout + c4 * 4);
#pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            acc[px].x = fmaf(x[px + dx], w.x, acc[px].x); acc[px].y = fmaf(x[px + dx], w.y, acc[px].y);
                            acc[px].z = fmaf(x[px + dx], w.z, acc[px].z); acc[px].w = fmaf(x[px + dx], w.w, acc[px].w);
                        }
                    }
                }
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                if (f0 + px < F) {
                    float4 a = acc[px];
                    a.x = fmaxf(a.x, 0.0f); a.y = fmaxf(a.y, 0.0f); a.z = fmaxf(a.z, 0.0f); a.w = fmaxf(a.w, 0.0f);
                    *reinterpret_cast<float4 *>(out + (((size_t)n * L + t) * F + f0 + px) * Cout + c4 * 4) = a;
                }
            }
        }
    }
}

// y [N][H][W][C] (GEMM result) -> out [N][ceil(H/2)][ceil(W/2)][C] = maxpool2x2_ceil( mask(relu(y + bias)) )
__global__ void __launch_bounds__(256)
bias_relu_mask_pool_kernel(const float *__restrict__ y, const float *__restrict__ bias, const int *__restrict__ valid,
                           int N, int H, int W, int C, float *__restrict__ out)
{
    const int C4 = C >> 2, H2 = (H + 1) >> 1, W2 = (W + 1) >> 1;
    const long long total = (long long)N * H2 * W2 * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long r = i / C4;
        const int w2 = (int)(r % W2); r /= W2;
        const int h2 = (int)(r % H2);
        const int n = (int)(r / H2);
        const int vr = __ldg(valid + n);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + c4);
        float4 m = make_float4(0.0f, 0.0f, 0.0f, 0.0f);              // relu output >= 0 and masked rows are 0
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int h = 2 * h2 + dy;
            if (h >= H || h >= vr) continue;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int w = 2 * w2 + dx;
                if (w >= W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4 *>(y + (((long long)n * H + h) * W + w) * C) + c4);
                m.x = fmaxf(m.x, __fadd_rn(v.x, b.x)); m.y = fmaxf(m.y, __fadd_rn(v.y, b.y));
                m.z = fmaxf(m.z, __fadd_rn(v.z, b.z)); m.w = fmaxf(m.w, __fadd_rn(v.w, b.w));
            }
        }
        reinterpret_cast<float4 *>(out)[i] = m;
    }
}

}  // namespace e2e

extern "C" int e2e_conv1_direct(const float *feat, long long feat_pitch_n, const float *weight, const float *bias,
                                const int *valid_rows, int N, int L, int F, int Cin, int Cout, float *out_nhwc, void *stream)
{
    using namespace e2e;
    if (!feat || !weight || !bias || !valid_rows || !out_nhwc) return set_error(E2E_ERR_ARG, "e2e_conv1_direct: null pointer");
    if (N <= 0 || L <= 0 || F <= 0 || Cin <= 0 || Cout <= 0 || (Cout & 3) || N > 65535 || feat_pitch_n < (long long)L * Cin * F)
        return set_error(E2E_ERR_ARG, "e2e_conv1_direct: bad size (Cout must be a multiple of 4, N <= 65535)");
    if ((reinterpret_cast<uintptr_t>(bias) & 15) || (reinterpret_cast<uintptr_t>(out_nhwc) & 15))
        return set_error(E2E_ERR_ARG, "e2e_conv1_direct: misaligned buffer");
    const size_t smem = ((size_t)9 * Cin * Cout + (size_t)(kC1Rows + 2) * Cin * (F + 5) + 4) * 4;
    if (smem > 200 * 1024) return set_error(E2E_ERR_UNSUPPORTED, "e2e_conv1_direct: %zu bytes of shared memory needed", smem);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(conv1_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    conv1_direct_kernel<<<dim3((L + kC1Rows - 1) / kC1Rows, N), 256, smem, static_cast<cudaStream_t>(stream)>>>(feat, feat_pitch_n, weight, bias, valid_rows,
                                                                                      L, F, Cin, Cout, out_nhwc);
    count_launch();
    return check_launch("e2e_conv1_direct");
}

extern "C" int e2e_conv_bias_relu_mask_pool(const float *y_nhwc, const float *bias, const int *valid_rows, int N, int H, int W, int C,
                                            float *out_nhwc, void *stream)
{
    using namespace e2e;
    if (!y_nhwc || !bias || !valid_rows || !out_nhwc) return set_error(E2E_ERR_ARG, "e2e_conv_bias_relu_mask_pool: null pointer");
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return set_error(E2E_ERR_ARG, "e2e_conv_bias_relu_mask_pool: bad size");
    if ((reinterpret_cast<uintptr_t>(y_nhwc) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15) || (reinterpret_cast<uintptr_t>(out_nhwc) & 15))
        return set_error(E2E_ERR_ARG, "e2e_conv_bias_relu_mask_pool: misaligned buffer");
    const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    bias_relu_mask_pool_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(y_nhwc, bias, valid_rows, N, H, W, C, out_nhwc);
    count_launch();
    return check_launch("e2e_conv_bias_relu_mask_pool");
}
