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
