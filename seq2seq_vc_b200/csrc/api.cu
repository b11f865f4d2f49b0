// C-ABI plumbing: error strings, device check, launch counter, and the s2s_gemm dispatcher.
#include <atomic>
#include <cstring>

#include "common.cuh"

namespace s2s {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

char* last_error_buf() { return g_err; }
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int gemm_simt(const s2s_gemm_t& g, cudaStream_t st);
int gemm_tc(const s2s_gemm_t& g, cudaStream_t st);
int gemm_tc_split(const s2s_gemm_t& g, cudaStream_t st);
int gemm_tc_grouped(const s2s_gemm_t* gs, int n, cudaStream_t st);
size_t gemm_split_workspace_bytes(const s2s_gemm_t& g);

}  // namespace s2s

extern "C" const char* s2s_last_error(void) { return s2s::g_err; }
extern "C" int s2s_abi_version(void) { return S2S_ABI_VERSION; }
extern "C" int64_t s2s_launch_count(void) { return (int64_t)s2s::g_launches.load(); }

extern "C" int s2s_device_check(void) {
    int dev = 0, major = 0;
    S2S_CUDA_OK(cudaGetDevice(&dev));
    S2S_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return s2s::set_error(S2S_ERR_UNSUPPORTED, "device compute capability %d.x is not sm_100", major);
    return S2S_OK;
}

extern "C" int s2s_gemm(const s2s_gemm_t* g, int mode, void* stream) {
    S2S_REQUIRE(g != nullptr, "gemm: null descriptor");
    S2S_REQUIRE(g->A && g->B && g->C, "gemm: null operand pointer");
    S2S_REQUIRE(g->M >= 0 && g->N >= 0 && g->K >= 0 && g->taps >= 1 && g->batch1 >= 1 && g->batch2 >= 1,
                "gemm: bad shape M=%d N=%d K=%d taps=%d batch=%dx%d", g->M, g->N, g->K, g->taps, g->batch1, g->batch2);
    S2S_REQUIRE(g->a_rs == 1 || g->a_cs == 1, "gemm: A must be contiguous along m or k");
    S2S_REQUIRE(g->b_rs == 1 || g->b_cs == 1, "gemm: B must be contiguous along n or k");
    S2S_REQUIRE((long)g->batch1 * g->batch2 <= 65535, "gemm: too many batches");
    if (g->M == 0 || g->N == 0) return S2S_OK;
    if (mode == 0) return s2s::gemm_simt(*g, (cudaStream_t)stream);
    if (mode == 1) return s2s::gemm_tc(*g, (cudaStream_t)stream);
    if (mode == 2) return s2s::gemm_tc_split(*g, (cudaStream_t)stream);
    return s2s::set_error(S2S_ERR_INVALID, "gemm: unknown mode %d", mode);
}

extern "C" size_t s2s_gemm_workspace_bytes(const s2s_gemm_t* g) { return g ? s2s::gemm_split_workspace_bytes(*g) : 0; }

extern "C" int s2s_gemm_grouped(const s2s_gemm_t* gs, int n, int mode, void* stream) {
    S2S_REQUIRE(gs != nullptr && n >= 1, "gemm_grouped: bad arguments");
    if (mode == 1 && n > 1) {
        int done = 0;
        bool all = true;
        // full groups of up to 8 problems; anything the grouped kernel cannot take is launched on its own
        while (done < n && all) {
            const int m = n - done < 8 ? n - done : 8;
            const int rc = s2s::gemm_tc_grouped(gs + done, m, (cudaStream_t)stream);
            if (rc == S2S_OK) done += m;
            else if (rc == S2S_ERR_UNSUPPORTED) all = false;
            else return rc;
        }
        if (done == n) return S2S_OK;
        for (int i = done; i < n; ++i) {
            const int rc = s2s_gemm(gs + i, mode, stream);
            if (rc != S2S_OK) return rc;
        }
        return S2S_OK;
    }
    for (int i = 0; i < n; ++i) {
        const int rc = s2s_gemm(gs + i, mode, stream);
        if (rc != S2S_OK) return rc;
    }
    return S2S_OK;
}
