// Error string, launch accounting and the small size helpers of the C ABI.
#include "common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace e2e {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what)
{
    const cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) return E2E_OK;
    cudaGetLastError();   // clear the non-sticky launch error so the next call starts clean
    return set_error(E2E_ERR_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace e2e

extern "C" const char *e2e_last_error(void) { return e2e::g_err; }
extern "C" int e2e_abi_version(void) { return E2E_ABI_VERSION; }
extern "C" int e2e_padded_vocab(int V) { return V <= 0 ? 0 : (V + 3) & ~3; }
extern "C" long long e2e_launch_count(void) { return e2e::g_launches.load(std::memory_order_relaxed); }
// Kernels of this library that ran as nodes of a replayed CUDA graph (captured launches are counted at capture time only).
extern "C" void e2e_add_launch_count(long long n) { e2e::count_launch((int)n); }

// Host -> device copy of the valid frames of n rows of a zero-padded [U][Lmax][D] fp32 feature tensor (pinned host memory):
// row r of the DEVICE tensor receives the first lens[order[r]] frames of HOST row order[r], one cudaMemcpyAsync per row on
// `stream`.  After every `chunk` rows the event events[r / chunk] (cudaEvent_t, created by the caller) is recorded, so that
// a consumer stream can start on the first chunk while the rest is still in flight.  Plain C loop: ~2 us per copy instead
// of the ~8 us a Python-level tensor.copy_ costs (BeamDecoder.decode_batch_from_host; bin/test_asr.py:161-163 `.to(device)`).
extern "C" int e2e_copy_rows_h2d(const float *host, float *dev, long long row_pitch, int D, const int *lens_host,
                                 const long long *order_host, int n, int chunk, void *const *events, void *stream)
{
    using namespace e2e;
    if (!host || !dev || !lens_host || !order_host || n < 0 || D <= 0 || row_pitch <= 0)
        return set_error(E2E_ERR_ARG, "e2e_copy_rows_h2d: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int r = 0; r < n; ++r) {
        const long long u = order_host[r];
        const int len = lens_host[u];
        if (len > 0) {
            const cudaError_t e = cudaMemcpyAsync(dev + (long long)r * row_pitch, host + u * row_pitch, (size_t)len * D * sizeof(float),
                                                  cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "e2e_copy_rows_h2d: %s", cudaGetErrorString(e));
        }
        if (events && chunk > 0 && ((r + 1) % chunk == 0 || r + 1 == n)) {
            const cudaError_t e = cudaEventRecord(static_cast<cudaEvent_t>(events[r / chunk]), st);
            if (e != cudaSuccess) return set_error(E2E_ERR_LAUNCH, "e2e_copy_rows_h2d: %s", cudaGetErrorString(e));
        }
    }
    return E2E_OK;
}
