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
