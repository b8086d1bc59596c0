// Kernel (2): batched CTC prefix-score recursion.
// Replaces CTCPrefixScore.cheap_compute / full_compute (src/ctc.py:29-108) for every live
// (utterance, beam slot, candidate) in one launch.
//
// Mapping.  A prefix-state "lane" l = slot*C + j of one utterance walks the encoder frames
// t = start..T-1 sequentially in fp32 (the reference's order).  Per frame it needs
//     r0' = logaddexp(r0, phi) + x_c      r1' = logaddexp(r1, r0) + x_blank       (the recurrence)
//     psi = logaddexp(psi, phi + x_c)                                              (a running reduction)
// psi never feeds back into r, so every lane is served by TWO threads of the same CTA: a
// "state" thread carries (r0, r1) and streams them out as one coalesced float2 per frame, a
// "psi" thread carries psi.  That doubles the independent dependency chains in flight per
// utterance, which is what the sequential-in-T recursion is starved of.
// A CTA covers up to 128 consecutive lanes of ONE utterance (2x that many threads), so its
// threads share the utterance's posterior rows x[t][u][:] and the few parent states.
// Frames are processed in tiles of kTile:
//   * "rows" variant (Vp <= kMaxRowFloats): the tile's posterior rows are brought into shared
//     memory by the TMA engine (cp.async.bulk, one 16B-aligned row per copy, completion on an
//     mbarrier), double buffered, so the gather x[t][cand] becomes a conflict-free LDS;
//   * "gather" variant (large vocabularies): each state thread fetches its own column
//     x[t][u][cand] for the whole tile with independent loads and parks it in shared memory.
//   The parents' states are turned into phi tiles in shared memory once per hypothesis
//     phi[h][t] = ( logaddexp(r_prev[t][0], r_prev[t][1]),  r_prev[t][1] )
//   and are software pipelined: the global loads for tile k+1 are issued before tile k is
//   computed, so one __syncthreads per tile is all the synchronisation there is.
#include "common.cuh"

namespace e2e {

constexpr int kTile = 32;            // frames per shared-memory tile
constexpr int kPhiPitch = kTile + 1; // float2 pitch of a phi row (odd: no bank conflicts across hypotheses)
constexpr int kMaxRowFloats = 256;   // rows variant up to 1 KB per posterior row
constexpr int kMaxLanes = 128;       // lanes per CTA (threads = 2 x lanes)
constexpr int kPrefetch = 2;         // phi entries per thread held in registers across a tile
constexpr int kLutBytes = kLutNodes * kLutCopies * 16;

struct PrefixParams {
    const float *x; int Tmax, U, Vp, V;
    const int *enc_len;
    const float2 *r_prev; int lanes_prev;
    const int *prev_lane, *last_tok, *prefix_len, *n_live, *cand;
    int B, C, flags;
    float *psi; float2 *r_out; int *status;
    int chunks_per_utt;   // CTAs per utterance
    int lanes_per_cta;    // multiple of 32; blockDim.x = 2 * lanes_per_cta
    int hyps_per_cta;     // rows of the phi tile
};

__host__ __device__ inline size_t prefix_xs_bytes(bool gather, int lanes, int Vp)
{
    size_t f = gather ? (size_t)kTile * lanes + kTile : (size_t)2 * kTile * Vp;
    return ((f + 3) & ~(size_t)3) * 4;
}
__host__ __device__ inline size_t prefix_phis_bytes(int H) { return (size_t)2 * H * kPhiPitch * 8; }

template <bool kGather, int kMath>
__global__ void __launch_bounds__(2 * kMaxLanes, 5)
prefix_score_kernel(const PrefixParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nl = p.lanes_per_cta;
    const bool psi_role = tid >= nl;              // warp uniform (nl is a multiple of 32)
    const int lt = psi_role ? tid - nl : tid;     // lane index within the CTA
    const int u = blockIdx.x / p.chunks_per_utt;
    const int chunk = blockIdx.x % p.chunks_per_utt;
    const int T = p.enc_len ? p.enc_len[u] : p.Tmax;
    const int live = p.n_live ? p.n_live[u] : p.B;
    const int C = p.C, LU = p.B * p.C;
    const int lane0 = chunk * nl;                 // first lane (within the utterance) of this CTA
    if (T <= 0 || lane0 >= live * C) return;      // nothing to do for this CTA (uniform exit)
    const bool full = (p.flags & E2E_PREFIX_FULL) != 0;
    const bool fill_dead = (p.flags & E2E_PREFIX_SKIP_DEAD_ROWS) == 0;

    // ---- shared memory: LUT replicas | x tiles | phi tiles (x2) | s_plane[H] | s_red[2] | mbarriers[2]
    const int H = p.hyps_per_cta;
    const size_t xs_off = kLutBytes;
    const size_t phis_off = xs_off + prefix_xs_bytes(kGather, nl, p.Vp);
    const size_t misc_off = phis_off + prefix_phis_bytes(H);
    const size_t bars_off = (misc_off + (size_t)(H + 2) * 4 + 7) & ~(size_t)7;
    float4 *lut_base = reinterpret_cast<float4 *>(smem_raw);
    float *xs = reinterpret_cast<float *>(smem_raw + xs_off);
    float2 *phis = reinterpret_cast<float2 *>(smem_raw + phis_off);
    int *s_plane = reinterpret_cast<int *>(smem_raw + misc_off);
    int *s_red = s_plane + H;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + bars_off);
    const float4 *lut = lut_base + (tid & (kLutCopies - 1));      // this thread's replica
    float *xb_g = xs + (size_t)kTile * nl;                        // gather variant only

    // ---- per-lane setup (identical in both roles) ---------------------------------------------
    const int lane_u = lane0 + lt;                // lane within the utterance
    const bool active = lane_u < live * C;
    const int h = active ? lane_u / C : 0;        // beam slot
    const int j = active ? lane_u - h * C : 0;    // candidate index
    const int h_lo = lane0 / C;
    const int n = u * p.B + h;                    // hypothesis index
    int tok = 0, plen = 0, ltok = 0;
    if (active) {
        tok = full ? j : p.cand[(long long)n * C + j];
        plen = p.prefix_len[n];
        ltok = p.last_tok[n];
    }
    const int start = plen > 1 ? plen : 1;
    const bool too_long = active && (start - 1 >= T);
    if (too_long && !psi_role && p.status) atomicOr(p.status + u, E2E_STATUS_PREFIX_TOO_LONG);
    const bool run = active && !too_long;
    const bool special = full ? (tok == (plen > 0 ? ltok : 0)) : (plen > 0 && tok == ltok);

    if (tid == 0) s_red[0] = 0x7fffffff;
    if (kMath == kMathLut) softplus_lut_to_smem(lut_base, tid, nt);
    for (int i = tid; i < H; i += nt) {
        const int hh = h_lo + i;
        s_plane[i] = (hh < live) ? p.prev_lane[u * p.B + hh] : -1;
    }
    if (!kGather && tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (run && !psi_role) atomicMin(&s_red[0], start);
    __syncthreads();
    const int cta_start = s_red[0];               // 0x7fffffff if no lane runs
    const long long xrow0 = (long long)u * p.Vp;  // x[t][u][:] = x + t*U*Vp + xrow0
    const long long xstride = (long long)p.U * p.Vp;
    float2 *__restrict__ rout = p.r_out + ((long long)u * p.Tmax) * LU + lane_u;
    const float2 *__restrict__ rprev_u = p.r_prev + ((long long)u * p.Tmax) * p.lanes_prev;
    const float2 dead = make_float2(E2E_CTC_LOGZERO, E2E_CTC_LOGZERO);

    float nb = E2E_CTC_LOGZERO, bl = E2E_CTC_LOGZERO;
    if (run && plen == 0) nb = __ldg(p.x + xrow0 + tok);     // r[0,0,:] = x[0, c]  (src/ctc.py:82-83)
    float psi = nb;                                            // psi = r[start-1, 0, :] (src/ctc.py:85)

    if (cta_start != 0x7fffffff) {
        const int first_tile = cta_start / kTile;
        const int n_tiles = (T + kTile - 1) / kTile;
        // rows below the first computed tile are log-zero by construction
        if (run && !psi_role && fill_dead)
            for (int t = 0; t < first_tile * kTile && t < T; ++t) rout[(long long)t * LU] = dead;

        auto issue_rows = [&](int k) {   // warp 0: one bulk copy per posterior row of tile k
            const int t0 = k * kTile;
            const int rows = min(kTile, T - t0);
            uint64_t *bar = &bars[k & 1];
            float *dst = xs + (size_t)(k & 1) * kTile * p.Vp;
            if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)rows * p.Vp * 4u);
            __syncwarp();
            if (tid < rows)
                bulk_g2s(dst + (size_t)tid * p.Vp, p.x + (long long)(t0 + tid) * xstride + xrow0, (uint32_t)p.Vp * 4u, bar);
        };
        // phi tile k, entry (hl, tt) describes r_prev at frame k*kTile + tt - 1
        const int n_phi = H * kTile;
        auto phi_fetch = [&](int k, int i) -> float2 {
            const int hl = i / kTile, tt = i - hl * kTile;
            const int ts = k * kTile + tt - 1;
            const int pl = s_plane[hl];
            return (pl >= 0 && ts >= 0 && ts < T) ? __ldg(rprev_u + (long long)ts * p.lanes_prev + pl) : dead;
        };
        auto phi_put = [&](int k, int i, float2 a) {
            const int hl = i / kTile, tt = i - hl * kTile;
            float2 ph;
            ph.x = logaddexp<kMath>(a.x, a.y, lut);
            ph.y = full ? logaddexp<kMath>(E2E_CTC_LOGZERO, a.y, lut) : a.y;
            phis[(size_t)(k & 1) * H * kPhiPitch + hl * kPhiPitch + tt] = ph;
        };
        auto phi_load = [&](int k, float2 (&reg)[kPrefetch]) {
#pragma unroll
            for (int q = 0; q < kPrefetch; ++q) {
                const int i = tid + q * nt;
                if (i < n_phi) reg[q] = phi_fetch(k, i);
            }
        };
        auto phi_store = [&](int k, const float2 (&reg)[kPrefetch]) {
#pragma unroll
            for (int q = 0; q < kPrefetch; ++q) {
                const int i = tid + q * nt;
                if (i < n_phi) phi_put(k, i, reg[q]);
            }
            for (int i = tid + kPrefetch * nt; i < n_phi; i += nt) phi_put(k, i, phi_fetch(k, i));   // unusual shapes only
        };

        float2 pf[kPrefetch];
#pragma unroll
        for (int q = 0; q < kPrefetch; ++q) pf[q] = dead;
        if (first_tile < n_tiles) {
            if (!kGather && tid < 32) {
                issue_rows(first_tile);
                if (first_tile + 1 < n_tiles) issue_rows(first_tile + 1);
            }
            phi_load(first_tile, pf);
            phi_store(first_tile, pf);
        }

        uint32_t parity[2] = {0u, 0u};
        for (int k = first_tile; k < n_tiles; ++k) {
            const int t0 = k * kTile;
            const int rows = min(kTile, T - t0);
            if (k + 1 < n_tiles) phi_load(k + 1, pf);          // in flight while tile k is computed
            const float *xt;   // tile base: xt[tt*pitch + col]
            int pitch, col;
            if (kGather) {
                __syncthreads();                               // everyone is done with the previous tile's columns
                if (active && !psi_role) {
#pragma unroll 8
                    for (int tt = 0; tt < rows; ++tt)
                        xs[tt * nl + lt] = __ldg(p.x + (long long)(t0 + tt) * xstride + xrow0 + tok);
                }
                if (tid < rows) xb_g[tid] = __ldg(p.x + (long long)(t0 + tid) * xstride + xrow0 + E2E_CTC_BLANK);
                xt = xs; pitch = nl; col = lt;
            } else {
                mbar_wait(&bars[k & 1], parity[k & 1]);
                parity[k & 1] ^= 1u;
                xt = xs + (size_t)(k & 1) * kTile * p.Vp; pitch = p.Vp; col = tok;
            }
            __syncthreads();    // phi tile k and x tile k visible; all threads have left tile k-1
            if (!kGather && tid < 32 && k > first_tile && k + 1 < n_tiles) issue_rows(k + 1);   // reuses tile k-1's buffer
            if (run) {
                const float2 *ph_row = phis + (size_t)(k & 1) * H * kPhiPitch + (size_t)(h - h_lo) * kPhiPitch;
                int tt = 0;
                const int tt_first = start - t0;              // first frame of this tile that is computed
                if (tt_first > 0) {
                    const int stop = min(tt_first, rows);
                    if (!psi_role) {
                        // Row 0 of an empty-prefix extension, (x[0,c], logzero), is read back by the child's
                        // first frame (its start is also 1), so it is written even when dead rows are skipped.
                        if (t0 == 0 && plen == 0) rout[0] = make_float2(nb, E2E_CTC_LOGZERO);
                        if (fill_dead)
                            for (; tt < stop; ++tt)
                                if (!(t0 + tt == 0 && plen == 0)) rout[(long long)(t0 + tt) * LU] = dead;
                    }
                    tt = stop;
                }
                const float2 *php = ph_row + tt;
                const float *xcp = xt + tt * pitch + col;
                if (!psi_role) {
                    // state thread: 3 LDS, 2 log-add-exp, 1 STG.64 per frame
                    const float *xbp = kGather ? (xb_g + tt) : (xt + tt * pitch + E2E_CTC_BLANK);
                    const int xb_step = kGather ? 1 : pitch;
                    float2 *outp = rout + (long long)(t0 + tt) * LU;
#pragma unroll 4
                    for (; tt < rows; ++tt) {
                        const float2 ph = *php;
                        const float phi = special ? ph.y : ph.x;
                        const float nnb = __fadd_rn(logaddexp<kMath>(nb, phi, lut), *xcp);
                        const float nbl = __fadd_rn(logaddexp<kMath>(bl, nb, lut), *xbp);
                        nb = nnb; bl = nbl;
                        *outp = make_float2(nnb, nbl);
                        ++php; xcp += pitch; xbp += xb_step; outp += LU;
                    }
                } else {
                    // psi thread: 2 LDS, 1 log-add-exp per frame
#pragma unroll 4
                    for (; tt < rows; ++tt) {
                        const float2 ph = *php;
                        const float phi = special ? ph.y : ph.x;
                        psi = logaddexp<kMath>(psi, __fadd_rn(phi, *xcp), lut);
                        ++php; xcp += pitch;
                    }
                }
            }
            if (k + 1 < n_tiles) phi_store(k + 1, pf);         // other phi buffer: nobody reads it before the next barrier
        }
    }

    if (run) {
        const bool eos_lane = !full && tok == E2E_CTC_EOS;       // P(<eos> | g) = P(g)   (src/ctc.py:106-107)
        if (eos_lane && (psi_role || (start >= T && fill_dead))) {
            const float2 a = __ldg(rprev_u + (long long)(T - 1) * p.lanes_prev + s_plane[h - h_lo]);
            psi = logaddexp<kMath>(a.x, a.y, lut);
            // psi aliases r[start-1,0,:] in the reference when the time loop never runs (src/ctc.py:85)
            if (!psi_role) rout[(long long)(start - 1) * LU] = make_float2(psi, E2E_CTC_LOGZERO);
        }
        if (psi_role) p.psi[(long long)n * C + j] = psi;
    } else if (active && psi_role) {
        p.psi[(long long)n * C + j] = E2E_CTC_LOGZERO;
    }
}

static size_t prefix_smem_bytes(bool gather, int lanes, int Vp, int H)
{
    size_t b = kLutBytes + prefix_xs_bytes(gather, lanes, Vp) + prefix_phis_bytes(H);
    b = ((b + (size_t)(H + 2) * 4 + 7) & ~(size_t)7) + 16;
    return (b + 15) & ~(size_t)15;
}

}  // namespace e2e

extern "C" int e2e_ctc_prefix_score(const float *x, int Tmax, int U, int Vp, int V, const int *enc_len,
                                    const float *r_prev, int lanes_prev,
                                    const int *prev_lane, const int *last_tok, const int *prefix_len,
                                    const int *n_live, const int *cand, int B, int C, int flags,
                                    float *psi, float *r_out, int *status, int n_run, void *stream)
{
    using namespace e2e;
    const bool full = (flags & E2E_PREFIX_FULL) != 0;
    if (!x || !r_prev || !prev_lane || !last_tok || !prefix_len || !psi || !r_out || (!full && !cand))
        return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_score: null pointer");
    if (Tmax <= 0 || U <= 0 || V <= 0 || B <= 0 || C <= 0 || lanes_prev <= 0)
        return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_score: non-positive size");
    if (Vp != e2e_padded_vocab(V)) return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_score: Vp=%d, expected %d", Vp, e2e_padded_vocab(V));
    if (full && C != V) return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_score: E2E_PREFIX_FULL needs C == V");
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(r_out) & 7) || (reinterpret_cast<uintptr_t>(r_prev) & 7))
        return set_error(E2E_ERR_ARG, "e2e_ctc_prefix_score: misaligned buffer");

    const long long LU = (long long)B * C;
    int nl = (int)((LU + 31) / 32 * 32);
    if (nl > kMaxLanes) {
        // split the utterance's lanes over several CTAs of equal, warp-multiple size
        const int chunks = (int)((LU + kMaxLanes - 1) / kMaxLanes);
        nl = (int)(((LU + chunks - 1) / chunks + 31) / 32 * 32);
    }
    PrefixParams p;
    p.x = x; p.Tmax = Tmax; p.U = U; p.Vp = Vp; p.V = V; p.enc_len = enc_len;
    p.r_prev = reinterpret_cast<const float2 *>(r_prev); p.lanes_prev = lanes_prev;
    p.prev_lane = prev_lane; p.last_tok = last_tok; p.prefix_len = prefix_len; p.n_live = n_live; p.cand = cand;
    p.B = B; p.C = C; p.flags = flags;
    p.psi = psi; p.r_out = reinterpret_cast<float2 *>(r_out); p.status = status;
    p.lanes_per_cta = nl;
    p.chunks_per_utt = (int)((LU + nl - 1) / nl);
    p.hyps_per_cta = (nl + C - 1) / C + 1;
    if (p.hyps_per_cta > B) p.hyps_per_cta = B;
    if (n_run <= 0 || n_run > U) n_run = U;      // only the first n_run utterances are processed
    const long long grid = (long long)n_run * p.chunks_per_utt;
    if (grid > 0x7fffffffLL) return set_error(E2E_ERR_UNSUPPORTED, "e2e_ctc_prefix_score: grid too large");

    const bool gather = Vp > kMaxRowFloats;
    const int math = (flags & E2E_PREFIX_FAST_MATH) ? kMathMufu : ((flags & E2E_PREFIX_LIBM_MATH) ? kMathLibm : kMathLut);
    const size_t smem = prefix_smem_bytes(gather, nl, Vp, p.hyps_per_cta);
    void (*kern)(PrefixParams);
    if (gather)
        kern = math == kMathLut ? prefix_score_kernel<true, kMathLut>
                                : (math == kMathMufu ? prefix_score_kernel<true, kMathMufu> : prefix_score_kernel<true, kMathLibm>);
    else
        kern = math == kMathLut ? prefix_score_kernel<false, kMathLut>
                                : (math == kMathMufu ? prefix_score_kernel<false, kMathMufu> : prefix_score_kernel<false, kMathLibm>);
    if (smem > 48 * 1024) {
        if (smem > 200 * 1024) return set_error(E2E_ERR_UNSUPPORTED, "e2e_ctc_prefix_score: %zu bytes of shared memory needed", smem);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    kern<<<(unsigned)grid, 2 * nl, smem, static_cast<cudaStream_t>(stream)>>>(p);
    count_launch();
    return check_launch("e2e_ctc_prefix_score");
}
