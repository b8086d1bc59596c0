/* Plain-C restatement of the CTC prefix-score recursion (TEST INFRASTRUCTURE).
 *
 * Not part of the product.  Restates /root/reference/src/ctc.py:
 *   oracle_blank_state   <- CTCPrefixScore.init_state      (ctc.py:19-27)
 *   oracle_extend        <- CTCPrefixScore.cheap_compute   (ctc.py:68-108)
 *                           and .full_compute (full_mode=1) (ctc.py:29-66)
 * and numpy's fp32 logaddexp loop (the arithmetic the reference leans on:
 *   x==y -> x+ln2 ; d=x-y ; d>0 -> x+log1pf(expf(-d)) ; else y+log1pf(expf(d))).
 * Checked bit-for-bit against oracle/ctc_prefix_oracle.py (and through it the
 * live reference) in tests/test_oracle_c.py.  Built by oracle/Makefile with
 * -O2 -ffp-contract=off (no fused multiply-add, no fast-math).
 *
 * oracle_extend_many runs independent extend() calls over pthreads; it is
 * the kernel-level CPU baseline of bench.py (kind "port").
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <pthread.h>
#include <unistd.h>

#define ORACLE_LOGZERO (-100000000.0f)
#define ORACLE_BLANK 0
#define ORACLE_EOS 1
#define ORACLE_LN2F 0.693147180559945309417232121458176568F

static inline float lae(float a, float b)
{
    if (a == b) return a + ORACLE_LN2F;
    {
        const float d = a - b;
        if (d > 0) return a + log1pf(expf(-d));
        if (d <= 0) return b + log1pf(expf(d));
        return d; /* NaN */
    }
}

float oracle_logaddexpf(float a, float b) { return lae(a, b); }

/* x: [T, ldx] rows of V log-posteriors; r: [T,2] */
void oracle_blank_state(const float *x, int T, int ldx, float *r)
{
    float acc = x[ORACLE_BLANK];
    int t;
    r[0] = ORACLE_LOGZERO;
    r[1] = acc;
    for (t = 1; t < T; ++t) {
        acc = acc + x[(size_t)t * ldx + ORACLE_BLANK];
        r[2 * t] = ORACLE_LOGZERO;
        r[2 * t + 1] = acc;
    }
}

/* Returns 0, or -1 when the prefix is longer than T (the reference raises
 * IndexError at ctc.py:85).  psi: [C]; r: [C, T, 2] contiguous.
 * full_mode: candidates are 0..V-1 (cands may be NULL, C must equal V), the
 * last-token column uses logaddexp(logzero, r_prev[t,1]) and there is no eos
 * override (ctc.py:53-56,65). */
int oracle_extend(const float *x, int T, int V, int ldx,
                  int prefix_len, int last_tok, const float *r_prev,
                  const int *cands, int C, int full_mode,
                  float *psi, float *r)
{
    const int start = prefix_len > 1 ? prefix_len : 1;
    int j, t;
    (void)V;
    if (start - 1 >= T) return -1;
    for (j = 0; j < C; ++j) {
        const int c = full_mode ? j : cands[j];
        float *rj = r + (size_t)j * T * 2;
        float nb, bl, p;
        int special;
        for (t = 0; t < 2 * T; ++t) rj[t] = ORACLE_LOGZERO;
        if (prefix_len == 0) rj[0] = x[c];
        if (full_mode) special = (c == (prefix_len > 0 ? last_tok : 0));
        else special = (prefix_len > 0 && c == last_tok);
        nb = rj[2 * (start - 1)];
        bl = rj[2 * (start - 1) + 1];
        p = nb;
        for (t = start; t < T; ++t) {
            const float a0 = r_prev[2 * (t - 1)], a1 = r_prev[2 * (t - 1) + 1];
            float phi;
            const float xc = x[(size_t)t * ldx + c];
            const float xb = x[(size_t)t * ldx + ORACLE_BLANK];
            float nnb, nbl;
            if (special) phi = full_mode ? lae(ORACLE_LOGZERO, a1) : a1;
            else phi = lae(a0, a1);
            nnb = lae(nb, phi) + xc;
            nbl = lae(bl, nb) + xb;
            p = lae(p, phi + xc);
            rj[2 * t] = nnb;
            rj[2 * t + 1] = nbl;
            nb = nnb;
            bl = nbl;
        }
        if (!full_mode && c == ORACLE_EOS) {
            p = lae(r_prev[2 * (T - 1)], r_prev[2 * (T - 1) + 1]);
            if (start >= T) rj[2 * (start - 1)] = p; /* view-aliasing quirk, ctc.py:85,107 */
        }
        psi[j] = p;
    }
    return 0;
}

/* n independent calls sharing one posterior matrix layout:
 *   call i uses x + x_off[i] (T[i] rows, stride ldx), r_prev + rp_off[i],
 *   cands + i*C, writes psi + i*C and r + r_off[i].
 * Work is handed out in chunks of 4 calls to `threads` pthreads (threads<=0 ->
 * one per online core).  Returns the number of failed calls. */
typedef struct {
    int n;
    const float *x; const long long *x_off; const int *T; int V; int ldx;
    const int *prefix_len; const int *last_tok;
    const float *r_prev; const long long *rp_off;
    const int *cands; int C;
    float *psi; float *r; const long long *r_off;
    int next; int bad;
    pthread_mutex_t mu;
} many_job;

static void *many_worker(void *arg)
{
    many_job *jb = (many_job *)arg;
    for (;;) {
        int lo, hi, i, bad = 0;
        pthread_mutex_lock(&jb->mu);
        lo = jb->next;
        jb->next += 4;
        pthread_mutex_unlock(&jb->mu);
        if (lo >= jb->n) break;
        hi = lo + 4 < jb->n ? lo + 4 : jb->n;
        for (i = lo; i < hi; ++i) {
            int rc = oracle_extend(jb->x + jb->x_off[i], jb->T[i], jb->V, jb->ldx, jb->prefix_len[i],
                                   jb->last_tok[i], jb->r_prev + jb->rp_off[i],
                                   jb->cands + (size_t)i * jb->C, jb->C, 0,
                                   jb->psi + (size_t)i * jb->C, jb->r + jb->r_off[i]);
            if (rc) bad += 1;
        }
        if (bad) {
            pthread_mutex_lock(&jb->mu);
            jb->bad += bad;
            pthread_mutex_unlock(&jb->mu);
        }
    }
    return NULL;
}

int oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int oracle_extend_many(int n, const float *x, const long long *x_off, const int *T, int V, int ldx,
                       const int *prefix_len, const int *last_tok,
                       const float *r_prev, const long long *rp_off,
                       const int *cands, int C,
                       float *psi, float *r, const long long *r_off, int threads)
{
    many_job jb;
    pthread_t tid[256];
    int k, started = 0;
    if (threads <= 0) threads = oracle_max_threads();
    if (threads > 256) threads = 256;
    jb.n = n; jb.x = x; jb.x_off = x_off; jb.T = T; jb.V = V; jb.ldx = ldx;
    jb.prefix_len = prefix_len; jb.last_tok = last_tok; jb.r_prev = r_prev; jb.rp_off = rp_off;
    jb.cands = cands; jb.C = C; jb.psi = psi; jb.r = r; jb.r_off = r_off;
    jb.next = 0; jb.bad = 0;
    pthread_mutex_init(&jb.mu, NULL);
    for (k = 1; k < threads; ++k) {
        if (pthread_create(&tid[started], NULL, many_worker, &jb) != 0) break;
        ++started;
    }
    many_worker(&jb);
    for (k = 0; k < started; ++k) pthread_join(tid[k], NULL);
    pthread_mutex_destroy(&jb.mu);
    return jb.bad;
}
