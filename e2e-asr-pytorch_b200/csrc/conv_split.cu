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
//
// Second operand format (kPieces == 2, see lstm_step.cu): two fp16 pieces of scale*x — three partial GEMM products
// instead of six and a third less unfolded data.  The activations of the front end are ReLU outputs of no fixed
// range, so the scale is chosen ON THE DEVICE per layer input: the kernel that produces a layer's input
// (conv1_direct / the bias-ReLU-mask epilogues) also reduces max|x| into a device word (float bits; x >= 0, so the
// bit patterns order like the values), the unfold kernel turns it into the power of two that puts that maximum in
// [2^14, 2^15), and the epilogue of the layer's own GEMM removes it again (exactly).  No host round trip.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace e2e {

// amax bits (a non-negative float) -> scale = 2^(14 - floor(log2(amax))) as float bits, clamped to 2^-60 .. 2^60;
// amax == 0 (an all-zero block) -> 1
__device__ __forceinline__ float act_scale_from_amax(unsigned bits)
{
    const int e = (int)((bits >> 23) & 0xffu);                // biased exponent of amax
    if (bits == 0u || e == 0) return 1.0f;                    // zero / subnormal maximum
    int se = 127 + 14 - (e - 127);                            // biased exponent of the scale
    se = se < 127 - 60 ? 127 - 60 : (se > 127 + 60 ? 127 + 60 : se);
    return __uint_as_float((unsigned)se << 23);
}

// max over the CTA's threads of a non-negative value, one atomicMax per warp
__device__ __forceinline__ void amax_reduce_store(float v, unsigned *amax_out)
{
    unsigned b = __float_as_uint(v) & 0x7fffffffu;          // a -0.0 must not look like the largest word
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned other = __shfl_xor_sync(0xffffffffu, b, o);
        b = other > b ? other : b;
    }
    if ((threadIdx.x & 31) == 0 && b != 0u) atomicMax(amax_out, b);
}

struct alignas(8) ConvF16x4 { __half v[4]; };

__device__ __forceinline__ void conv_split3(float x, __nv_bfloat16 &a1, __nv_bfloat16 &a2, __nv_bfloat16 &a3)
{
    a1 = __float2bfloat16_rn(x);
    const float r1 = __fsub_rn(x, __bfloat162float(a1));
    a2 = __float2bfloat16_rn(r1);
    a3 = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(a2)));
}

struct alignas(8) ConvBf16x4 { __nv_bfloat16 v[4]; };

// in  [N][H][W][C] fp32 (NHWC), valid [N] rows per image; pixels p0 .. p0+P-1 of the flattened (n,h,w) index
// out [P][kPieces*9C] 16-bit pieces (kPieces == 3: bf16; kPieces == 2: fp16 of scale*x, scale from *amax)
template <int kPieces>
__global__ void __launch_bounds__(256)
im2col3x3_split_kernel(const float *__restrict__ in, const int *__restrict__ valid, int H, int W, int C,
                       long long p0, int P, unsigned short *__restrict__ out, const unsigned *__restrict__ amax)
{
    // one warp per output pixel (its coordinates are decoded once, warp uniform); the lanes sweep the
    // pixel's 9 taps x C/4 channel groups: 512-byte contiguous reads, 256-byte contiguous writes per piece
    const int C4 = C >> 2;
    const long long K = 9LL * C;
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    float scale = 1.0f;
    if (kPieces == 2) scale = act_scale_from_amax(__ldg(amax));
    for (int pl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pl < P; pl += warps_per_grid) {
        const long long p = p0 + pl;
        const int w = (int)(p % W);
        const long long nh = p / W;
        const int h = (int)(nh % H);
        const int n = (int)(nh / H);
        const int vrows = __ldg(valid + n);
        const float *img = in + (long long)n * H * W * C;
        unsigned short *orow = out + (long long)pl * kPieces * K;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int hs = h + dy - 1;
            const bool row_ok = hs >= 0 && hs < H && hs < vrows;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ws = w + dx - 1;
                const bool ok = row_ok && ws >= 0 && ws < W;
                const float4 *src = reinterpret_cast<const float4 *>(img + ((long long)hs * W + ws) * C);
                unsigned short *o = orow + (dy * 3 + dx) * C;
                for (int c4 = lane; c4 < C4; c4 += 32) {
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (ok) v = __ldg(src + c4);
                    if (kPieces == 3) {
                        ConvBf16x4 q0, q1, q2;
                        conv_split3(v.x, q0.v[0], q1.v[0], q2.v[0]);
                        conv_split3(v.y, q0.v[1], q1.v[1], q2.v[1]);
                        conv_split3(v.z, q0.v[2], q1.v[2], q2.v[2]);
                        conv_split3(v.w, q0.v[3], q1.v[3], q2.v[3]);
                        *reinterpret_cast<ConvBf16x4 *>(o + c4 * 4) = q0;
                        *reinterpret_cast<ConvBf16x4 *>(o + K + c4 * 4) = q1;
                        *reinterpret_cast<ConvBf16x4 *>(o + 2 * K + c4 * 4) = q2;
                    } else {
                        const float x[4] = {v.x, v.y, v.z, v.w};
                        ConvF16x4 q0, q1;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float xs = __fmul_rn(x[i], scale);
                            q0.v[i] = __float2half_rn(xs);
                            q1.v[i] = __float2half_rn(__fsub_rn(xs, __half2float(q0.v[i])));
                        }
                        *reinterpret_cast<ConvF16x4 *>(o + c4 * 4) = q0;
                        *reinterpret_cast<ConvF16x4 *>(o + K + c4 * 4) = q1;
                    }
                }
            }
        }
    }
}

// y[p][c] = relu(y[p][c] + bias[c]) for valid rows, 0 for rows h >= valid[n]   (in place, NHWC)
// kScaled: y is first multiplied by inv_w_scale / act_scale(*amax_in) — the power-of-two factor an fp16x2 GEMM result
// carries — and the maximum of the result is reduced into *amax_out (the next layer's scale), if given.
template <bool kScaled>
__global__ void __launch_bounds__(256)
bias_relu_mask_kernel(float *__restrict__ y, const float *__restrict__ bias, const int *__restrict__ valid,
                      int H, int W, int C, long long p0, long long P,
                      const unsigned *__restrict__ amax_in, float inv_w_scale, unsigned *__restrict__ amax_out)
{
    const int C4 = C >> 2;
    const long long total = P * C4;
    float gs = 1.0f, vmax = 0.0f;
    if (kScaled) gs = __fdiv_rn(inv_w_scale, act_scale_from_amax(__ldg(amax_in)));
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
            if (kScaled) { v.x = __fmul_rn(v.x, gs); v.y = __fmul_rn(v.y, gs); v.z = __fmul_rn(v.z, gs); v.w = __fmul_rn(v.w, gs); }
            v.x = fmaxf(__fadd_rn(v.x, b.x), 0.0f); v.y = fmaxf(__fadd_rn(v.y, b.y), 0.0f);
            v.z = fmaxf(__fadd_rn(v.z, b.z), 0.0f); v.w = fmaxf(__fadd_rn(v.w, b.w), 0.0f);
            if (kScaled) vmax = fmaxf(fmaxf(vmax, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        } else {
            v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        *ptr = v;
    }
    if (kScaled && amax_out) amax_reduce_store(vmax, amax_out);
}

}  // namespace e2e

namespace {

using namespace e2e;

int unfold_impl(const char *name, int pieces, const float *in_nhwc, const int *valid_rows, int N, int H, int W, int C,
                long long first_pixel, int n_pixels, const unsigned *amax_in, void *out16, void *stream)
{
    if (!in_nhwc || !valid_rows || !out16 || (pieces == 2 && !amax_in)) return set_error(E2E_ERR_ARG, "%s: null pointer", name);
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || n_pixels <= 0 || first_pixel < 0 ||
        first_pixel + n_pixels > (long long)N * H * W)
        return set_error(E2E_ERR_ARG, "%s: bad size (C must be a multiple of 4)", name);
    if ((reinterpret_cast<uintptr_t>(in_nhwc) & 15) || (reinterpret_cast<uintptr_t>(out16) & 7))
        return set_error(E2E_ERR_ARG, "%s: misaligned buffer", name);
    long long blocks = ((long long)n_pixels + 7) / 8;           // 8 warps per block, one pixel per warp at a time
    if (blocks > 148LL * 64) blocks = 148LL * 64;             // grid-stride: a few waves of the machine
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned short *out = static_cast<unsigned short *>(out16);
    if (pieces == 3)
        im2col3x3_split_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(in_nhwc, valid_rows, H, W, C, first_pixel, n_pixels, out, nullptr);
    else
        im2col3x3_split_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(in_nhwc, valid_rows, H, W, C, first_pixel, n_pixels, out, amax_in);
    count_launch();
    return check_launch(name);
}

int bias_relu_mask_impl(const char *name, bool scaled, float *y_nhwc, const float *bias, const int *valid_rows, int N, int H, int W, int C,
                        long long first_pixel, long long n_pixels, const unsigned *amax_in, float inv_w_scale, unsigned *amax_out,
                        void *stream)
{
    if (!y_nhwc || !bias || !valid_rows || (scaled && !amax_in)) return set_error(E2E_ERR_ARG, "%s: null pointer", name);
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3) || n_pixels <= 0 || first_pixel < 0 ||
        first_pixel + n_pixels > (long long)N * H * W)
        return set_error(E2E_ERR_ARG, "%s: bad size (C must be a multiple of 4)", name);
    if (scaled && !(inv_w_scale > 0.0f)) return set_error(E2E_ERR_ARG, "%s: inv_w_scale must be positive", name);
    if ((reinterpret_cast<uintptr_t>(y_nhwc) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15))
        return set_error(E2E_ERR_ARG, "%s: misaligned buffer", name);
    const long long total = n_pixels * (C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (scaled)
        bias_relu_mask_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(y_nhwc, bias, valid_rows, H, W, C, first_pixel, n_pixels,
                                                                      amax_in, inv_w_scale, amax_out);
    else
        bias_relu_mask_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(y_nhwc, bias, valid_rows, H, W, C, first_pixel, n_pixels,
                                                                       nullptr, 1.0f, nullptr);
    count_launch();
    return check_launch(name);
}

}  // namespace

extern "C" int e2e_conv3x3_unfold_split(const float *in_nhwc, const int *valid_rows, int N, int H, int W, int C,
                                        long long first_pixel, int n_pixels, void *out_bf16, void *stream)
{
    return unfold_impl("e2e_conv3x3_unfold_split", 3, in_nhwc, valid_rows, N, H, W, C, first_pixel, n_pixels, nullptr, out_bf16, stream);
}

extern "C" int e2e_conv3x3_unfold_split_f16x2(const float *in_nhwc, const int *valid_rows, int N, int H, int W, int C,
                                              long long first_pixel, int n_pixels, const unsigned *amax_in, void *out_f16, void *stream)
{
    return unfold_impl("e2e_conv3x3_unfold_split_f16x2", 2, in_nhwc, valid_rows, N, H, W, C, first_pixel, n_pixels, amax_in, out_f16, stream);
}

extern "C" int e2e_conv_bias_relu_mask(float *y_nhwc, const float *bias, const int *valid_rows, int N, int H, int W, int C,
                                       long long first_pixel, long long n_pixels, void *stream)
{
    return bias_relu_mask_impl("e2e_conv_bias_relu_mask", false, y_nhwc, bias, valid_rows, N, H, W, C, first_pixel, n_pixels,
                               nullptr, 1.0f, nullptr, stream);
}

extern "C" int e2e_conv_bias_relu_mask_scaled(float *y_nhwc, const float *bias, const int *valid_rows, int N, int H, int W, int C,
                                              long long first_pixel, long long n_pixels, const unsigned *amax_in, float inv_w_scale,
                                              unsigned *amax_out, void *stream)
{
    return bias_relu_mask_impl("e2e_conv_bias_relu_mask_scaled", true, y_nhwc, bias, valid_rows, N, H, W, C, first_pixel, n_pixels,
                               amax_in, inv_w_scale, amax_out, stream);
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

// kTrack: also reduce the maximum of the outputs on rows t < valid[n] (the rows the next layer reads) into *amax_out.
template <bool kTrack>
__global__ void __launch_bounds__(256)
conv1_direct_kernel(const float *__restrict__ feat, long long feat_pitch_n, const float *__restrict__ weight,
                    const float *__restrict__ bias, const int *__restrict__ valid, int L, int F, int Cin, int Cout,
                    float *__restrict__ out, unsigned *__restrict__ amax_out)
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
    float vmax = 0.0f;
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
                    if (kTrack && t < vr) vmax = fmaxf(fmaxf(vmax, fmaxf(a.x, a.y)), fmaxf(a.z, a.w));
                }
            }
        }
    }
    if (kTrack) amax_reduce_store(vmax, amax_out);
}

// y [N][H][W][C] (GEMM result) -> out [N][ceil(H/2)][ceil(W/2)][C] = maxpool2x2_ceil( mask(relu(y + bias)) )
// kScaled as in bias_relu_mask_kernel.
template <bool kScaled>
__global__ void __launch_bounds__(256)
bias_relu_mask_pool_kernel(const float *__restrict__ y, const float *__restrict__ bias, const int *__restrict__ valid,
                           int N, int H, int W, int C, float *__restrict__ out,
                           const unsigned *__restrict__ amax_in, float inv_w_scale, unsigned *__restrict__ amax_out)
{
    const int C4 = C >> 2, H2 = (H + 1) >> 1, W2 = (W + 1) >> 1;
    const long long total = (long long)N * H2 * W2 * C4;
    float gs = 1.0f, vmax = 0.0f;
    if (kScaled) gs = __fdiv_rn(inv_w_scale, act_scale_from_amax(__ldg(amax_in)));
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
                float4 v = __ldg(reinterpret_cast<const float4 *>(y + (((long long)n * H + h) * W + w) * C) + c4);
                if (kScaled) { v.x = __fmul_rn(v.x, gs); v.y = __fmul_rn(v.y, gs); v.z = __fmul_rn(v.z, gs); v.w = __fmul_rn(v.w, gs); }
                m.x = fmaxf(m.x, __fadd_rn(v.x, b.x)); m.y = fmaxf(m.y, __fadd_rn(v.y, b.y));
                m.z = fmaxf(m.z, __fadd_rn(v.z, b.z)); m.w = fmaxf(m.w, __fadd_rn(v.w, b.w));
            }
        }
        reinterpret_cast<float4 *>(out)[i] = m;
        if (kScaled) vmax = fmaxf(fmaxf(vmax, fmaxf(m.x, m.y)), fmaxf(m.z, m.w));
    }
    if (kScaled && amax_out) amax_reduce_store(vmax, amax_out);
}

}  // namespace e2e

namespace {

int conv1_direct_impl(const char *name, const float *feat, long long feat_pitch_n, const float *weight, const float *bias,
                      const int *valid_rows, int N, int L, int F, int Cin, int Cout, float *out_nhwc, unsigned *amax_out, void *stream)
{
    if (!feat || !weight || !bias || !valid_rows || !out_nhwc) return set_error(E2E_ERR_ARG, "%s: null pointer", name);
    if (N <= 0 || L <= 0 || F <= 0 || Cin <= 0 || Cout <= 0 || (Cout & 3) || N > 65535 || feat_pitch_n < (long long)L * Cin * F)
        return set_error(E2E_ERR_ARG, "%s: bad size (Cout must be a multiple of 4, N <= 65535)", name);
    if ((reinterpret_cast<uintptr_t>(bias) & 15) || (reinterpret_cast<uintptr_t>(out_nhwc) & 15))
        return set_error(E2E_ERR_ARG, "%s: misaligned buffer", name);
    const size_t smem = ((size_t)9 * Cin * Cout + (size_t)(kC1Rows + 2) * Cin * (F + 5) + 4) * 4;
    if (smem > 200 * 1024) return set_error(E2E_ERR_UNSUPPORTED, "%s: %zu bytes of shared memory needed", name, smem);
    auto kern = amax_out ? conv1_direct_kernel<true> : conv1_direct_kernel<false>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<dim3((L + kC1Rows - 1) / kC1Rows, N), 256, smem, static_cast<cudaStream_t>(stream)>>>(feat, feat_pitch_n, weight, bias, valid_rows,
                                                                                       L, F, Cin, Cout, out_nhwc, amax_out);
    count_launch();
    return check_launch(name);
}

int bias_relu_mask_pool_impl(const char *name, bool scaled, const float *y_nhwc, const float *bias, const int *valid_rows,
                             int N, int H, int W, int C, float *out_nhwc, const unsigned *amax_in, float inv_w_scale,
                             unsigned *amax_out, void *stream)
{
    if (!y_nhwc || !bias || !valid_rows || !out_nhwc || (scaled && !amax_in)) return set_error(E2E_ERR_ARG, "%s: null pointer", name);
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 3)) return set_error(E2E_ERR_ARG, "%s: bad size", name);
    if (scaled && !(inv_w_scale > 0.0f)) return set_error(E2E_ERR_ARG, "%s: inv_w_scale must be positive", name);
    if ((reinterpret_cast<uintptr_t>(y_nhwc) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15) || (reinterpret_cast<uintptr_t>(out_nhwc) & 15))
        return set_error(E2E_ERR_ARG, "%s: misaligned buffer", name);
    const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (scaled)
        bias_relu_mask_pool_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(y_nhwc, bias, valid_rows, N, H, W, C, out_nhwc,
                                                                           amax_in, inv_w_scale, amax_out);
    else
        bias_relu_mask_pool_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(y_nhwc, bias, valid_rows, N, H, W, C, out_nhwc,
                                                                            nullptr, 1.0f, nullptr);
    count_launch();
    return check_launch(name);
}

}  // namespace

extern "C" int e2e_conv1_direct(const float *feat, long long feat_pitch_n, const float *weight, const float *bias,
                                const int *valid_rows, int N, int L, int F, int Cin, int Cout, float *out_nhwc, void *stream)
{
    return conv1_direct_impl("e2e_conv1_direct", feat, feat_pitch_n, weight, bias, valid_rows, N, L, F, Cin, Cout, out_nhwc, nullptr, stream);
}

extern "C" int e2e_conv1_direct_amax(const float *feat, long long feat_pitch_n, const float *weight, const float *bias,
                                     const int *valid_rows, int N, int L, int F, int Cin, int Cout, float *out_nhwc,
                                     unsigned *amax_out, void *stream)
{
    if (!amax_out) return e2e::set_error(E2E_ERR_ARG, "e2e_conv1_direct_amax: null pointer");
    return conv1_direct_impl("e2e_conv1_direct_amax", feat, feat_pitch_n, weight, bias, valid_rows, N, L, F, Cin, Cout, out_nhwc, amax_out, stream);
}

extern "C" int e2e_conv_bias_relu_mask_pool(const float *y_nhwc, const float *bias, const int *valid_rows, int N, int H, int W, int C,
                                            float *out_nhwc, void *stream)
{
    return bias_relu_mask_pool_impl("e2e_conv_bias_relu_mask_pool", false, y_nhwc, bias, valid_rows, N, H, W, C, out_nhwc,
                                    nullptr, 1.0f, nullptr, stream);
}

extern "C" int e2e_conv_bias_relu_mask_pool_scaled(const float *y_nhwc, const float *bias, const int *valid_rows, int N, int H, int W, int C,
                                                   float *out_nhwc, const unsigned *amax_in, float inv_w_scale, unsigned *amax_out,
                                                   void *stream)
{
    return bias_relu_mask_pool_impl("e2e_conv_bias_relu_mask_pool_scaled", true, y_nhwc, bias, valid_rows, N, H, W, C, out_nhwc,
                                    amax_in, inv_w_scale, amax_out, stream);
}
