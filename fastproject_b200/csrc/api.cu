// Library-level entry points: version, error text, launch counter, and the
// precision-mode dispatch of fp_consistency.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "fp_common.cuh"

int fp_consistency_fp64(const double *coords, const double *sigma_sq, double sigma_sq_scalar,
                        const double *ranks, int64_t N, int64_t K, int64_t ld, double *out_dissim,
                        int64_t ld_out, int output_kind, cudaStream_t st);

size_t fp_consistency_tc_workspace_bytes(int64_t N, int64_t K);
int fp_consistency_tc_status(const void *workspace, int64_t N, int64_t K, cudaStream_t st);
size_t fp_values_pack_tc_bytes(int64_t N, int64_t K);
int fp_values_pack_tc(const double *values, int64_t K, int64_t N, int64_t ld, void *packed, size_t packed_bytes,
                      cudaStream_t st);
int fp_consistency_tc(const double *coords, const double *sigma_sq, double sigma_sq_scalar,
                      const double *values, const void *packed, int64_t N, int64_t K, int64_t ld, double *out,
                      int64_t ld_out, int output_kind, void *workspace, size_t workspace_bytes, cudaStream_t st);

namespace fp {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
void clear_error() { g_error[0] = 0; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

}  // namespace fp

extern "C" int fp_abi_version(void) { return FP_ABI_VERSION; }
extern "C" const char *fp_version(void) { return "fastproject_b200 0.1.0 (sm_100a)"; }
extern "C" const char *fp_last_error(void) { return fp::g_error; }
extern "C" uint64_t fp_launch_count(void) { return fp::g_launches.load(); }

extern "C" size_t fp_consistency_workspace_bytes(int64_t N, int64_t K, int precision_mode) {
    if (precision_mode == FP_PRECISION_TC) return fp_consistency_tc_workspace_bytes(N, K);
    return 0;
}

extern "C" int fp_consistency(const double *coords, const double *sigma_sq, double sigma_sq_scalar,
                              const double *ranks, int64_t N, int64_t K, int64_t ld,
                              double *out_dissim, int64_t ld_out, int output_kind,
                              int precision_mode, void *workspace, size_t workspace_bytes, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(coords && ranks && out_dissim, "fp_consistency: null pointer");
    FP_CHECK_ARG(N > 0 && K >= 0 && ld >= N && ld_out >= N,
                 "fp_consistency: bad shape N=%ld K=%ld ld=%ld ld_out=%ld", (long)N, (long)K, (long)ld,
                 (long)ld_out);
    FP_CHECK_ARG(sigma_sq || sigma_sq_scalar > 0.0, "fp_consistency: sigma_sq must be positive");
    FP_CHECK_ARG((output_kind & ~FP_OUT_ANY_CELL_ORDER) == FP_OUT_ABS_ERROR ||
                     (output_kind & ~FP_OUT_ANY_CELL_ORDER) == FP_OUT_PREDICTION,
                 "fp_consistency: unknown output kind %d", output_kind);
    if (K == 0) return FP_OK;
    if (precision_mode == FP_PRECISION_FP64)
        return fp_consistency_fp64(coords, sigma_sq, sigma_sq_scalar, ranks, N, K, ld, out_dissim,
                                   ld_out, output_kind & ~FP_OUT_ANY_CELL_ORDER, fp::as_stream(stream));
    if (precision_mode == FP_PRECISION_TC)
        return fp_consistency_tc(coords, sigma_sq, sigma_sq_scalar, ranks, nullptr, N, K, ld, out_dissim, ld_out,
                                 output_kind, workspace, workspace_bytes, fp::as_stream(stream));
    fp::set_error("fp_consistency: unknown precision mode %d", precision_mode);
    return FP_EUNSUPPORTED;
}

extern "C" size_t fp_values_pack_bytes(int64_t N, int64_t K) { return fp_values_pack_tc_bytes(N, K); }

extern "C" int fp_values_pack(const double *values, int64_t K, int64_t N, int64_t ld, void *packed, size_t packed_bytes,
                              void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(values && packed, "fp_values_pack: null pointer");
    FP_CHECK_ARG(N > 0 && K > 0 && ld >= N, "fp_values_pack: bad shape N=%ld K=%ld ld=%ld", (long)N, (long)K, (long)ld);
    return fp_values_pack_tc(values, K, N, ld, packed, packed_bytes, fp::as_stream(stream));
}

extern "C" int fp_consistency_packed(const double *coords, const double *sigma_sq, double sigma_sq_scalar,
                                     const void *packed, int64_t N, int64_t K, double *out_dissim, int64_t ld_out,
                                     int output_kind, void *workspace, size_t workspace_bytes, void *stream) {
    fp::clear_error();
    FP_CHECK_ARG(coords && packed && out_dissim, "fp_consistency_packed: null pointer");
    FP_CHECK_ARG(N > 0 && K > 0 && ld_out >= N, "fp_consistency_packed: bad shape N=%ld K=%ld ld_out=%ld", (long)N,
                 (long)K, (long)ld_out);
    FP_CHECK_ARG(sigma_sq || sigma_sq_scalar > 0.0, "fp_consistency_packed: sigma_sq must be positive");
    FP_CHECK_ARG((output_kind & ~FP_OUT_ANY_CELL_ORDER) == FP_OUT_ABS_ERROR ||
                     (output_kind & ~FP_OUT_ANY_CELL_ORDER) == FP_OUT_PREDICTION,
                 "fp_consistency_packed: unknown output kind %d", output_kind);
    return fp_consistency_tc(coords, sigma_sq, sigma_sq_scalar, nullptr, packed, N, K, 0, out_dissim, ld_out,
                             output_kind, workspace, workspace_bytes, fp::as_stream(stream));
}

namespace {
__global__ void __launch_bounds__(256) pull_host_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16,
                                                        size_t tail_bytes) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
    if (blockIdx.x == 0 && threadIdx.x < tail_bytes)
        reinterpret_cast<unsigned char *>(dst + n16)[threadIdx.x] =
            reinterpret_cast<const unsigned char *>(src + n16)[threadIdx.x];
}
}  // namespace

extern "C" int fp_pull_host(void *dst_device, const void *src_pinned_host, size_t bytes, void *stream) {
    fp::clear_error();
    if (bytes == 0) return FP_OK;
    FP_CHECK_ARG(dst_device && src_pinned_host, "fp_pull_host: null pointer");
    FP_CHECK_ARG(((reinterpret_cast<uintptr_t>(dst_device) | reinterpret_cast<uintptr_t>(src_pinned_host)) & 15) == 0,
                 "fp_pull_host: pointers must be 16-byte aligned");
    const size_t n16 = bytes / 16;
    size_t blocks = (n16 + 255) / 256;
    const size_t cap = (size_t)fp::sm_count() * 4;       // many loads in flight: PCIe reads are latency bound
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pull_host_kernel<<<(unsigned)blocks, 256, 0, fp::as_stream(stream)>>>(
        static_cast<uint4 *>(dst_device), static_cast<const uint4 *>(src_pinned_host), n16, bytes % 16);
    FP_CHECK_LAUNCH("pull_host_kernel");
    return FP_OK;
}

extern "C" int fp_consistency_status(const void *workspace, int64_t N, int64_t K, int precision_mode,
                                     void *stream) {
    fp::clear_error();
    if (precision_mode != FP_PRECISION_TC) {
        FP_CHECK_CUDA(cudaStreamSynchronize(fp::as_stream(stream)));
        return FP_OK;
    }
    FP_CHECK_ARG(workspace && N > 0 && K > 0, "fp_consistency_status: bad arguments");
    return fp_consistency_tc_status(workspace, N, K, fp::as_stream(stream));
}
