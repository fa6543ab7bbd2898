// frontend.cu — SURVEY.md 8(f) row 1, first half: the duplicate filter of the reference's front end
// (Graph::initialize, graph.cpp:273-303 + 340-361) on the device.  The reference runs it as an O(group^2) nested loop
// on one core (20 s of a 5.9 M-overlap repeat-heavy probe, SURVEY.md finding 7); here every record decides for itself
// by scanning its own query group (common.cuh duplicate_filter_keeps): the three 4-byte columns of a group sit in a
// few L1 lines, neighbouring threads scan the same lines, and nothing is written but one byte per record.
// Not part of the graph step (the drop-in's initialize() is unchanged host code): a stateless stage on host buffers,
// like rala_b200_trim_classify.
#include "kernels.h"
#include "session.h"

namespace rb {

__global__ void __launch_bounds__(256) k_filter_duplicates(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                           const uint32_t* __restrict__ len, uint32_t n, uint8_t* __restrict__ valid) {
    for (uint32_t k0 = blockIdx.x * blockDim.x; k0 < n; k0 += gridDim.x * blockDim.x) {
        const uint32_t k = k0 + threadIdx.x;
        if (k < n) valid[k] = duplicate_filter_keeps(a, b, len, n, k) ? 1 : 0;
    }
}

void launch_filter_duplicates(Launch& L, const uint32_t* a, const uint32_t* b, const uint32_t* len, uint32_t n, uint8_t* valid) {
    if (n == 0) return;
    const uint64_t blocks = ((uint64_t) n + 255) / 256;
    k_filter_duplicates<<<(unsigned) (blocks < (1u << 20) ? blocks : (1u << 20)), 256, 0, L.stream>>>(a, b, len, n, valid);
    L.count++;
}

void preload_frontend() {
    cudaFuncAttributes at;
    cudaFuncGetAttributes(&at, k_filter_duplicates);
}

}  // namespace rb

using namespace rb;

extern "C" int rala_b200_filter_duplicates(rala_b200_ctx* ctx, const uint32_t* a_id, const uint32_t* b_id, const uint32_t* length,
                                           uint64_t n, uint8_t* valid_out, float* device_ms) {
    if (!ctx || (n && (!a_id || !b_id || !length || !valid_out))) return RALA_B200_ERR_ARG;
    if (n >= (1ull << 31)) return fail(ctx, RALA_B200_ERR_LIMIT, "too many overlap records (%llu >= 2^31)", (unsigned long long) n);
    if (device_ms) *device_ms = 0.f;
    if (n == 0) return RALA_B200_OK;
    CU(ctx, cudaSetDevice(ctx->device));
    DevBuf cols, valid;
    const size_t col = align_up((size_t) n * 4, 256);
    CU(ctx, cols.reserve(3 * col));
    CU(ctx, valid.reserve(align_up((size_t) n, 256)));
    char* base = cols.as<char>();
    cudaStream_t s = ctx->L.stream;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // the context's own pair belongs to rala_b200_event_record
    auto cleanup = [&](int rc) {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        cols.release();
        valid.release();
        return rc;
    };
#define CUF(call)                                                                                                        \
    do {                                                                                                                 \
        cudaError_t err__ = (call);                                                                                      \
        if (err__ != cudaSuccess)                                                                                        \
            return cleanup(fail(ctx, RALA_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, __LINE__)); \
    } while (0)
    CUF(cudaEventCreate(&ev0));
    CUF(cudaEventCreate(&ev1));
    CUF(cudaMemcpyAsync(base, a_id, (size_t) n * 4, cudaMemcpyHostToDevice, s));
    CUF(cudaMemcpyAsync(base + col, b_id, (size_t) n * 4, cudaMemcpyHostToDevice, s));
    CUF(cudaMemcpyAsync(base + 2 * col, length, (size_t) n * 4, cudaMemcpyHostToDevice, s));
    CUF(cudaEventRecord(ev0, s));
    launch_filter_duplicates(ctx->L, (const uint32_t*) base, (const uint32_t*) (base + col), (const uint32_t*) (base + 2 * col), (uint32_t) n,
                             valid.as<uint8_t>());
    CUF(cudaGetLastError());
    CUF(cudaEventRecord(ev1, s));
    CUF(cudaMemcpyAsync(valid_out, valid.p, (size_t) n, cudaMemcpyDeviceToHost, s));
    CUF(cudaStreamSynchronize(s));
    if (device_ms) CUF(cudaEventElapsedTime(device_ms, ev0, ev1));
#undef CUF
    return cleanup(RALA_B200_OK);
}
