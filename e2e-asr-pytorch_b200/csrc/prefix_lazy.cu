// Kernel (2), batched beam-search form: ONE fused launch per decode step that
//   (a) brings the prefix state r of every LIVE hypothesis up to date (B chains per utterance), and
//   (b) scores all B*C candidate extensions (psi only).
// Replaces CTCPrefixScore.cheap_compute (src/ctc.py:68-108) as src/decode.py:131 calls it, with LAZY state
// evaluation: psi of an extension only needs the state of its PARENT (src/ctc.py:103), and of the
// B*C states cheap_compute builds only the <= B the beam keeps are ever read again (src/decode.py:250-254).  So the
// state of a hypothesis is computed once, one step later, when it has survived the prune — by the same sequential
// fp32 recurrence over the same frames, hence the same bits — and the B*C candidate lanes run the psi reduction alone:
//     state lanes (B per utterance)  r0' = logaddexp(r0, phi_parent) + x[t][tok]    r1' = logaddexp(r1, r0) + x[t][blank]
//     psi lanes (B*C per utterance)  psi = logaddexp(psi, phi_hyp[t-1] + x[t][cand])
// Per utterance-frame that is 2B + B*C log-add-exp chains steps (+ B for phi) instead of 3*B*C, and 8B bytes of state
// written instead of 8*B*C (12x less at C = 12).  The byte figure of SURVEY.md §8d (12 + 12/C per candidate-frame)
// stays the unit the roofline is reported in; the DRAM traffic actually moved is measured beside it.
//
// Mapping.  One CTA per utterance.  Warps 0..SW-1 are STATE warps (8 hypotheses each, 4 lanes per hypothesis:
// lane role 0 carries the r0 chain, role 1 the r1 chain one frame behind it — r1[t] needs r0[t-1], which travels by a
// shuffle issued a whole iteration before it is consumed, so no shuffle latency sits on the chain), the other PW warps
// are PSI warps (one lane per (hypothesis, candidate)).  Frames are processed in tiles of kT; state warps work one
// tile AHEAD of the psi warps and hand their tile over through shared memory; one __syncthreads per tile.
//   * posterior rows x[t][u][:] of a tile: one TMA box copy (cp.async.bulk.tensor, mbarrier completion), ring of 3;
//     large vocabularies (Vp > 256): every lane fetches its own column with 4-byte cp.async (LDGSTS), double buffered;
//   * the parents' states (previous step's buffer, lane = parent slot): 8-byte cp.async into a staging tile, turned into
//     phi_parent = logaddexp(r0,r1) (or r1 alone for a repeated token) by the state warp once per tile;
//   * the hypotheses' own new states: written to the step's output buffer [U][Tmax][B][2] and to a shared tile from
//     which every psi warp makes the (sum, blank) pairs of the <= 4 hypotheses its lanes belong to.
// All live hypotheses of an utterance have the same length s (the decode step), so the first frame is uniform per CTA.
// Step 0 (s == 0): the hypothesis is the empty prefix, whose state is the init buffer itself (passed as r_prev).
#include "common.cuh"
#include <string.h>
#include <cuda.h>

namespace e2e {

struct LazyParams {
    const float *x; int Tmax, U, Vp, V;
    const int *enc_len;
    const float2 *r_prev; int lanes_prev;
    const int *parent_slot, *last_tok, *parent_tok, *prefix_len, *n_live, *cand;
    int B, C, flags;
    float *psi; float2 *r_out; int *status;
    int state_warps, psi_warps, hyps_per_warp;
};

constexpr int kLazyMaxRowFloats = 256;

__host__ __device__ constexpr int lazy_conv_pitch(int tile) { return 2 * tile + 2; }

struct LazySmem {
    size_t xs, cur, pstage, pphi, conv, misc, bars, total;
};
__host__ __device__ inline LazySmem lazy_smem_layout(bool gather, int math, int threads, int Vp, int B, int SW, int PW, int NHW, int tile)
{
    LazySmem s;
    size_t off = (math == kMathLut) ? (size_t)kLutNodes * kLutCopies * 16 : 0;
    off = (off + 127) & ~(size_t)127;
    s.xs = off;
    off += gather ? (size_t)2 * tile * threads * 4 : (size_t)3 * tile * Vp * 4;
    off = (off + 127) & ~(size_t)127;
    s.cur = off;      off += (size_t)2 * (tile + 1) * B * 8;
    s.pstage = off;   off += (size_t)SW * 2 * 8 * tile * 8;
    s.pphi = off;     off += (size_t)SW * 8 * (tile + 1) * 4;
    s.conv = off;     off += (size_t)PW * NHW * lazy_conv_pitch(tile) * 4;
    s.misc = off;     off += (size_t)(2 * B) * 4;
    off = (off + 7) & ~(size_t)7;
    s.bars = off;     off += 3 * 8;
    s.total = (off + 15) & ~(size_t)15;
    return s;
}

__device__ __forceinline__ void lazy_tma_load_tile(void *dst_smem, const CUtensorMap *map, int u, int t0, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(0), "r"(u), "r"(t0), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void lazy_cp_async_8(void *dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void lazy_cp_async_4(void *dst_smem, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void lazy_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lazy_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef E2E_LAZY_MINBLOCKS
#define E2E_LAZY_MINBLOCKS 5
#endif

// kGather: column-gather staging of x (large vocabularies).  kVp: compile-time Vp (0 = from the parameters).
// kT: frames per tile.  kBig: CTAs of up to 1024 threads (beam sizes whose B*C lanes need more than 4 warps).
template <bool kGather, int kMath, int kVp, int kT, bool kBig>
__global__ void __launch_bounds__(kBig ? 1024 : 160, kBig ? 1 : E2E_LAZY_MINBLOCKS)
prefix_lazy_kernel(const LazyParams p, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int u = blockIdx.x;
    const int T = p.enc_len ? p.enc_len[u] : p.Tmax;
    const int live = p.n_live ? p.n_live[u] : p.B;
    const int B = p.B, C = p.C, SW = p.state_warps, PW = p.psi_warps, NHW = p.hyps_per_warp;
    const int Vp = kVp ? kVp : p.Vp;
    if (T <= 0 || live <= 0) return;                    // idle utterance (uniform exit)
    const int s = p.prefix_len[u * B];                  // length of every live hypothesis of this utterance
    const int start_h = s > 1 ? s : 1;                  // first frame of the psi reduction    (src/ctc.py:78)
    const int start_p = s > 2 ? s - 1 : 1;              // first frame of the state recurrence (the parents' start)
    const bool passthrough = s == 0;                    // the hypothesis IS the empty prefix: its state is r_prev
    constexpr int kConvP = lazy_conv_pitch(kT);

    const LazySmem L = lazy_smem_layout(kGather, kMath, nt, Vp, B, SW, PW, NHW, kT);
    float4 *lut_base = reinterpret_cast<float4 *>(smem_raw);
    float *xs = reinterpret_cast<float *>(smem_raw + L.xs);
    float2 *cur = reinterpret_cast<float2 *>(smem_raw + L.cur);
    float2 *pstage = reinterpret_cast<float2 *>(smem_raw + L.pstage);
    float *pphi = reinterpret_cast<float *>(smem_raw + L.pphi);
    float *conv = reinterpret_cast<float *>(smem_raw + L.conv);
    int *s_pslot = reinterpret_cast<int *>(smem_raw + L.misc);
    int *s_spec = s_pslot + B;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + L.bars);
    const uint32_t lut = softplus_lut_adj(lut_base + (tid & (kLutCopies - 1)));
    const float2 dead = make_float2(E2E_CTC_LOGZERO, E2E_CTC_LOGZERO);

    const bool is_state = wid < SW;
    // ---- too long: the reference raises IndexError at psi = r[start-1, 0, :] (src/ctc.py:85) ----------------------
    if (start_h - 1 >= T) {
        if (tid == 0 && p.status) atomicOr(p.status + u, E2E_STATUS_PREFIX_TOO_LONG);
        if (!is_state) {
            const int pl = (wid - SW) * 32 + lane;
            if (pl < live * C) p.psi[(long long)u * B * C + pl] = E2E_CTC_LOGZERO;
        }
        return;
    }

    // ---- per-lane setup ---------------------------------------------------------------------------------------------
    // state lanes
    const int hh = lane >> 2, role = lane & 3;
    const int hs = wid * 8 + hh;                                  // hypothesis (beam slot) of a state lane
    const bool s_act = is_state && hs < live && role < 2;
    int s_tok = 0;
    // psi lanes
    const int pw = wid - SW;
    const int pl = pw * 32 + lane;                                // lane within the utterance: h*C + j
    const bool p_act = !is_state && pl < live * C;
    const int ph = p_act ? pl / C : 0;
    const int pj = p_act ? pl - ph * C : 0;
    const int h_first = (pw * 32) / C;                            // first hypothesis this psi warp's lanes belong to
    int c_tok = 0;
    bool p_spec = false;
    if (is_state) {
        if (hs < live) {
            const int n = u * B + hs;
            s_tok = p.last_tok[n];
            if (role == 0) {
                s_pslot[hs] = p.parent_slot[n];
                s_spec[hs] = (s >= 2 && p.parent_tok && s_tok == p.parent_tok[n]) ? 1 : 0;    // repeated token (src/ctc.py:89-91)
            }
        } else if (hs < B && role == 0) {
            s_pslot[hs] = 0;
            s_spec[hs] = 0;
        }
    } else if (p_act) {
        const int n = u * B + ph;
        c_tok = p.cand[(long long)n * C + pj];
        p_spec = s > 0 && c_tok == p.last_tok[n];
    }
    if (kMath == kMathLut) softplus_lut_to_smem(lut_base, tid, nt);
    if (!kGather && tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const long long xrow0 = (long long)u * Vp;
    const long long xstride = (long long)p.U * Vp;
    const int jl = (T - 1) / kT;                                  // last tile
    const int j0s = start_p / kT;                                 // first tile of the state warps
    const int j0p = start_h / kT;                                 // first tile of the psi warps (j0s or j0s + 1)
    const float2 *__restrict__ rprev_u = p.r_prev + ((long long)u * p.Tmax) * p.lanes_prev;
    float2 *__restrict__ rout_u = p.r_out + ((long long)u * p.Tmax) * B;

    // posterior tile j -> ring stage j % 3 (rows variant)
    auto issue_x = [&](int j) {
        uint64_t *bar = &bars[j % 3];
        mbar_arrive_expect_tx(bar, (uint32_t)kT * Vp * 4u);
        lazy_tma_load_tile(xs + (size_t)(j % 3) * kT * Vp, &tmap, u, j * kT, bar);
    };
    auto wait_x = [&](int j) { mbar_wait(&bars[j % 3], (uint32_t)(((j - j0s) / 3) & 1)); };
    // gather variant: this thread's column of tile j -> its own slots of stage j & 1
    const int my_col = is_state ? (role == 0 ? s_tok : E2E_CTC_BLANK) : c_tok;
    auto fetch_col = [&](int j) {
        const int t0 = j * kT, rows = min(kT, T - t0);
        float *dst = xs + (size_t)(j & 1) * kT * nt + tid;
        const float *src = p.x + (long long)t0 * xstride + xrow0 + my_col;
        for (int tt = 0; tt < rows; ++tt) lazy_cp_async_4(dst + tt * nt, src + (long long)tt * xstride);
    };
    // raw parent states of tile j (entry (hh, tt) <-> frame j*kT + tt - 1) -> this state warp's staging slot j & 1
    float2 *my_stage = pstage + (size_t)wid * 2 * 8 * kT;
    float *my_pphi = pphi + (size_t)wid * 8 * (kT + 1);
    auto fetch_parents = [&](int j) {
        float2 *dst = my_stage + (size_t)(j & 1) * 8 * kT;
        for (int e = lane; e < 8 * kT; e += 32) {
            const int eh = e / kT, tt = e - eh * kT;
            const int ts = j * kT + tt - 1;
            const int hyp = wid * 8 + eh;
            if (hyp < live && ts >= 0 && ts < T) lazy_cp_async_8(dst + e, rprev_u + (long long)ts * p.lanes_prev + s_pslot[hyp]);
            else dst[e] = dead;
        }
    };

    if (!kGather && tid == 0) {
        issue_x(j0s);
        if (j0s + 1 <= jl) issue_x(j0s + 1);
    }
    if (is_state) {
        fetch_parents(j0s);
        if (kGather) fetch_col(j0s);
        lazy_cp_commit();
    } else if (kGather) {
        if (j0p <= jl) fetch_col(j0p);
        lazy_cp_commit();
    }

    // ---- chain registers ------------------------------------------------------------------------------------------
    // state lanes: v = r0 (role 0) / r1 (role 1, one frame behind); s0_prev = r0 two frames back (as role 1 needs it);
    // xl_prev = x[t-1][blank] for role 1.  The initial values make role 1's first (warm-up) step produce log-zero.
    float v = E2E_CTC_LOGZERO, s0_prev = E2E_CTC_LOGZERO, xl_prev = 0.0f;
    if (is_state && !passthrough && role == 0 && s == 1 && hs < live)
        v = __ldg(p.x + xrow0 + s_tok);                            // r[0,0] = x[0,c] for an extension of the empty prefix (src/ctc.py:82-83)
    float psi = E2E_CTC_LOGZERO;
    if (p_act && s == 0) psi = __ldg(p.x + xrow0 + c_tok);        // psi = r[start-1,0,:] (src/ctc.py:85)

    for (int k = j0s - 1; k <= jl; ++k) {
        if (!kGather && tid == 0 && k >= j0s && k + 2 <= jl) issue_x(k + 2);     // reuses the stage of tile k-1
        if (is_state) {
            const int j = k + 1;
            if (j <= jl) {
                const int t0 = j * kT, rows = min(kT, T - t0);
                float2 *cb = cur + (size_t)(j & 1) * (kT + 1) * B;
                if (j + 1 <= jl) {
                    fetch_parents(j + 1);
                    if (kGather) fetch_col(j + 1);
                }
                lazy_cp_commit();
                lazy_cp_wait<1>();                                  // tile j's parents (and column) have landed
                __syncwarp();
                const float2 *st = my_stage + (size_t)(j & 1) * 8 * kT;
                if (passthrough) {
                    // the empty prefix: its state rows are the parents' rows themselves
                    for (int e = lane; e < 8 * kT; e += 32) {
                        const int eh = e / kT, tt = e - eh * kT;
                        const int hyp = wid * 8 + eh;
                        if (hyp < B) cb[tt * B + hyp] = st[e];
                    }
                } else {
                    // phi of the parents for this tile: logaddexp(r0, r1), or r1 alone for a repeated token
                    for (int e = lane; e < 8 * kT; e += 32) {
                        const int eh = e / kT, tt = e - eh * kT;
                        const int hyp = wid * 8 + eh;
                        const float2 a = st[e];
                        const bool sp = hyp < B && s_spec[hyp] != 0;
                        my_pphi[eh * (kT + 1) + tt] = sp ? a.y : logaddexp<kMath>(a.x, a.y, lut);
                    }
                    __syncwarp();
                    if (!kGather) wait_x(j);
                    int tt = 0;
                    float *sp_ptr = reinterpret_cast<float *>(cb + hs) + (role == 0 ? 2 * B : 1);      // role 0: row tt+1 .x; role 1: row tt .y
                    float *gp_ptr = reinterpret_cast<float *>(rout_u + (long long)t0 * B + hs) + (role == 0 ? 0 : 1 - 2 * B);
                    if (j == j0s) {
                        tt = start_p - t0;
                        // the row before the first computed frame: (x[0,c] | log-zero, log-zero)
                        if (role == 0 && s_act) {
                            cb[tt * B + hs].x = v;
                            rout_u[(long long)(start_p - 1) * B + hs].x = v;
                        }
                    } else if (role == 0 && s_act) {
                        cb[hs].x = v;                               // r0 of the last frame of the previous tile
                    }
                    const float *php = my_pphi + hh * (kT + 1);
                    const float *xp;
                    int xstep;
                    if (kGather) { xp = xs + (size_t)(j & 1) * kT * nt + tid; xstep = nt; }
                    else { xp = xs + (size_t)(j % 3) * kT * Vp + (role == 0 ? s_tok : E2E_CTC_BLANK); xstep = Vp; }
                    const int src0 = lane & ~3;
                    auto frame = [&](int q) {
                        const float s0n = __shfl_sync(E2E_FULL_MASK, v, src0);
                        const float xl = xp[q * xstep];
                        const float b = role == 0 ? php[q] : s0_prev;
                        const float add = role == 0 ? xl : xl_prev;
                        v = __fadd_rn(logaddexp<kMath>(v, b, lut), add);
                        if (s_act) { sp_ptr[q * 2 * B] = v; gp_ptr[(long long)q * 2 * B] = v; }
                        s0_prev = s0n;
                        xl_prev = xl;
                    };
                    php += tt; xp += tt * xstep; sp_ptr += tt * 2 * B; gp_ptr += (long long)tt * 2 * B;
                    for (; tt + 4 <= rows; tt += 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) frame(q);
                        php += 4; xp += 4 * xstep; sp_ptr += 8 * B; gp_ptr += 8 * B;
                    }
                    for (; tt < rows; ++tt) {
                        frame(0);
                        php += 1; xp += xstep; sp_ptr += 2 * B; gp_ptr += 2 * B;
                    }
                    if (j == jl) {
                        // drain: r1 of the last frame (role 0 has nothing left to do)
                        const float r1_last = __fadd_rn(logaddexp<kMath>(v, s0_prev, lut), xl_prev);
                        if (role == 1 && s_act) { sp_ptr[0] = r1_last; gp_ptr[0] = r1_last; }
                    }
                }
            }
        } else {
            const int j = k;
            if (j >= j0p && j >= j0s) {
                const int t0 = j * kT, rows = min(kT, T - t0);
                const float2 *cb = cur + (size_t)(j & 1) * (kT + 1) * B;
                float *cw = conv + (size_t)pw * NHW * kConvP;
                if (kGather) {
                    if (j + 1 <= jl) fetch_col(j + 1);
                    lazy_cp_commit();
                    lazy_cp_wait<1>();
                }
                // (sum, blank) pairs of the hypotheses this warp's lanes belong to; entry tt <-> frame t0 + tt - 1
                for (int e = lane; e < NHW * kT; e += 32) {
                    const int hl = e / kT, tt = e - hl * kT;
                    const int hyp = h_first + hl;
                    if (hyp < B) {
                        const float2 a = cb[tt * B + hyp];
                        *reinterpret_cast<float2 *>(cw + hl * kConvP + 2 * tt) = make_float2(logaddexp<kMath>(a.x, a.y, lut), a.y);
                    }
                }
                __syncwarp();
                if (!kGather) wait_x(j);
                if (p_act) {
                    int tt = (j == j0p) ? start_h - t0 : 0;
                    const float *php = cw + (ph - h_first) * kConvP + (p_spec ? 1 : 0) + 2 * tt;
                    const float *xp;
                    int xstep;
                    if (kGather) { xp = xs + (size_t)(j & 1) * kT * nt + tid; xstep = nt; }
                    else { xp = xs + (size_t)(j % 3) * kT * Vp + c_tok; xstep = Vp; }
                    xp += tt * xstep;
                    for (; tt + 4 <= rows; tt += 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) psi = logaddexp<kMath>(psi, __fadd_rn(php[2 * q], xp[q * xstep]), lut);
                        php += 8; xp += 4 * xstep;
                    }
                    for (; tt < rows; ++tt) {
                        psi = logaddexp<kMath>(psi, __fadd_rn(php[0], xp[0]), lut);
                        php += 2; xp += xstep;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }

    if (p_act) {
        if (c_tok == E2E_CTC_EOS) {                                   // P(<eos> | g) = P(g)   (src/ctc.py:106-107)
            float2 a;
            if (passthrough) a = __ldg(rprev_u + (long long)(T - 1) * p.lanes_prev + s_pslot[ph]);
            else a = cur[(size_t)(jl & 1) * (kT + 1) * B + (T - 1 - jl * kT + 1) * B + ph];
            psi = logaddexp<kMath>(a.x, a.y, lut);
        }
        p.psi[(long long)(u * B + ph) * C + pj] = psi;
    }
}

typedef void (*LazyKernel)(LazyParams, CUtensorMap);

int make_posterior_map_lazy(CUtensorMap *map, const float *x, int Tmax, int U, int Vp, int tile);   // prefix_score.cu

template <bool kGather, int kVp, int kT, bool kBig>
static LazyKernel pick_math(int math)
{
    if (math == kMathPoly) return prefix_lazy_kernel<kGather, kMathPoly, kVp, kT, kBig>;
    if (math == kMathPolyEstrin) return prefix_lazy_kernel<kGather, kMathPolyEstrin, kVp, kT, kBig>;
    return prefix_lazy_kernel<kGather, kMathLut, kVp, kT, kBig>;
}

}  // namespace e2e

extern "C" int e2e_ctc_prefix_step_supported(int Vp, int B, int C)
{
    if (B <= 0 || C <= 0 || Vp <= 0) return 0;
    const int SW = (B + 7) / 8, PW = (B * C + 31) / 32;
    return (SW + PW) <= 32 ? 1 : 0;
}

extern "C" int e2e_ctc_prefix_step(const float *x, int Tmax, int U, int Vp, int V, const int *enc_len,
                                   const float *r_prev, int lanes_prev,
                                   const int *parent_slot, const int *last_tok, const int *parent_tok, const int *prefix_len,
                                   const int *n_live, const int *cand, int B, int C, int flags,
                                   float *psi, float *r_out, int *status, int n_run, void *stream)
{
    using namespace e2e;
    if (!x || !r_prev || !parent_slot || !last_tok || !prefix_len || !cand || !psi || !r_out)
        return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_step: null pointer");
    if (Tmax <= 0 || U <= 0 || V <= 0 || B <= 0 || C <= 0 || lanes_prev <= 0)
        return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_step: non-positive size");
    if (Vp != e2e_padded_vocab(V)) return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_step: Vp=%d, expected %d", Vp, e2e_padded_vocab(V));
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(r_out) & 7) || (reinterpret_cast<uintptr_t>(r_prev) & 7))
        return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_step: misaligned buffer");
    if (!e2e_ctc_prefix_step_supported(Vp, B, C))
        return set_error(E2E_ERR_UNSUPPORTED, "e2e_ctc_prefix_step: B=%d, C=%d needs more than 32 warps per utterance (use e2e_ctc_prefix_score)", B, C);
    if (n_run <= 0 || n_run > U) n_run = U;

    LazyParams p;
    p.x = x; p.Tmax = Tmax; p.U = U; p.Vp = Vp; p.V = V; p.enc_len = enc_len;
    p.r_prev = reinterpret_cast<const float2 *>(r_prev); p.lanes_prev = lanes_prev;
    p.parent_slot = parent_slot; p.last_tok = last_tok; p.parent_tok = parent_tok; p.prefix_len = prefix_len;
    p.n_live = n_live; p.cand = cand; p.B = B; p.C = C; p.flags = flags;
    p.psi = psi; p.r_out = reinterpret_cast<float2 *>(r_out); p.status = status;
    p.state_warps = (B + 7) / 8;
    p.psi_warps = (B * C + 31) / 32;
    p.hyps_per_warp = 31 / C + 2;
    if (p.hyps_per_warp > B) p.hyps_per_warp = B;
    const int threads = (p.state_warps + p.psi_warps) * 32;
    const bool big = threads > 160;
    const bool gather = Vp > kLazyMaxRowFloats;
    const int math = (flags & E2E_PREFIX_POLY_MATH) ? ((flags & E2E_PREFIX_POLY_ESTRIN) ? kMathPolyEstrin : kMathPoly) : kMathLut;
    if (flags & (E2E_PREFIX_FAST_MATH | E2E_PREFIX_LIBM_MATH | E2E_PREFIX_FULL))
        return set_error(E2E_ERR_UNSUPPORTED, "e2e_ctc_prefix_step: only the table and polynomial log-add-exp are built for the fused step");
    // 16-frame tiles for machine-filling launches (more CTAs per SM), 32-frame tiles for the tail (fewer barriers per chain)
    const bool small_tile = n_run >= 2 * 148;
    const int tile = small_tile ? 16 : 32;
    LazyKernel kern;
    if (gather) kern = small_tile ? (big ? pick_math<true, 0, 16, true>(math) : pick_math<true, 0, 16, false>(math))
                                  : (big ? pick_math<true, 0, 32, true>(math) : pick_math<true, 0, 32, false>(math));
    else if (Vp == 32) kern = small_tile ? (big ? pick_math<false, 32, 16, true>(math) : pick_math<false, 32, 16, false>(math))
                                         : (big ? pick_math<false, 32, 32, true>(math) : pick_math<false, 32, 32, false>(math));
    else kern = small_tile ? (big ? pick_math<false, 0, 16, true>(math) : pick_math<false, 0, 16, false>(math))
                           : (big ? pick_math<false, 0, 32, true>(math) : pick_math<false, 0, 32, false>(math));
    const LazySmem L = lazy_smem_layout(gather, math, threads, Vp, B, p.state_warps, p.psi_warps, p.hyps_per_warp, tile);
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (!gather) {
        const int rc = make_posterior_map_lazy(&map, x, Tmax, U, Vp, tile);
        if (rc != E2E_OK) return rc;
    }
    if (L.total > 48 * 1024) {
        if (L.total > 200 * 1024) return set_error(E2E_ERR_UNSUPPORTED, "e2e_ctc_prefix_step: %zu bytes of shared memory needed", L.total);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
        if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<(unsigned)n_run, threads, L.total, static_cast<cudaStream_t>(stream)>>>(p, map);
    count_launch();
    return check_launch("e2e_ctc_prefix_step");
}
