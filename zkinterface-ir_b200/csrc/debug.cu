// Debug / measurement entry points (not on the product path):
//   zkb_debug_field_ops         element-wise field ops on the device, both the PTX carry-chain product and the
//                               portable CIOS product, so tests can compare them with Python integers
//   zkb_debug_field_throughput  register-resident Montgomery products / modular additions per second: the
//                               integer-pipe ceiling of the level kernel
#include <algorithm>
#include <vector>

#include "context.h"
#include "device_util.cuh"

using namespace zkb;

#define CUDA_TRY(c, expr)                                                                              \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return (c)->fail(ZKB_E_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " #expr); \
    } while (0)

namespace {

// op: 0 add, 1 mont_mul (product path of the kernels), 2 mont_mul portable
template <int N>
__global__ void k_field_ops(int op, const uint32_t* a, const uint32_t* b, uint32_t* r, uint64_t n, FieldParams fp) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x[N], y[N], z[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        x[k] = a[i * N + k];
        y[k] = b[i * N + k];
    }
    if (op == 0) fe_add<N>(z, x, y, fp.p);
    else if (op == 1) fe_mont_mul<N>(z, x, y, fp.p, fp.n0inv);
    else fe_mont_mul_portable<N>(z, x, y, fp.p, fp.n0inv);
#pragma unroll
    for (int k = 0; k < N; k++) r[i * N + k] = z[k];
}

// a chain of dependent products per thread (x <- x*y), operands in registers: no memory traffic
template <int N>
__global__ void __launch_bounds__(256) k_field_throughput(int op, uint32_t iters, uint32_t* sink, FieldParams fp) {
    uint32_t x[N], y[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        x[k] = (threadIdx.x * 2654435761u + k * 40503u + blockIdx.x) & (k == N - 1 ? 0x0FFFFFFFu : 0xFFFFFFFFu);
        y[k] = (threadIdx.x * 97u + k * 7919u + 12345u) & (k == N - 1 ? 0x0FFFFFFFu : 0xFFFFFFFFu);
    }
    for (uint32_t i = 0; i < iters; i++) {
        uint32_t z[N];
        if (op == 0) fe_add<N>(z, x, y, fp.p);
        else fe_mont_mul<N>(z, x, y, fp.p, fp.n0inv);
#pragma unroll
        for (int k = 0; k < N; k++) x[k] = z[k];
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < N; k++) acc ^= x[k];
    if (acc == 0x12345678u) sink[0] = acc;  // keep the chain alive
}

}  // namespace

#define DISPATCH_N(nlimb, CALL)                        \
    switch (nlimb) {                                   \
        case 1: { constexpr int N = 1; CALL; } break;  \
        case 2: { constexpr int N = 2; CALL; } break;  \
        case 4: { constexpr int N = 4; CALL; } break;  \
        default: { constexpr int N = 8; CALL; } break; \
    }

extern "C" int zkb_debug_field_ops(zkb_ctx* c, int op, const uint32_t* a, const uint32_t* b, uint32_t* r, uint64_t n) {
    if (!c->prog.field_set || c->prog.binary) return c->fail(ZKB_E_ARG, "set an odd field first");
    if (!c->has_gpu) return c->fail(ZKB_E_CUDA, "no CUDA device in this context (there is no CPU fallback)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int N = c->prog.nlimb;
    size_t bytes = (size_t)n * N * 4;
    uint32_t *da = nullptr, *db = nullptr, *dr = nullptr;
    CUDA_TRY(c, cudaMalloc((void**)&da, bytes));
    CUDA_TRY(c, cudaMalloc((void**)&db, bytes));
    CUDA_TRY(c, cudaMalloc((void**)&dr, bytes));
    CUDA_TRY(c, cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, c->stream));
    const FieldParams fp = c->prog.fp;
    DISPATCH_N(N, (k_field_ops<N><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(op, da, db, dr, n, fp)));
    CUDA_TRY(c, cudaMemcpyAsync(r, dr, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dr);
    return ZKB_OK;
}

extern "C" int zkb_debug_field_throughput(zkb_ctx* c, int op, uint32_t iters, double* ops_per_second) {
    if (!c->prog.field_set || c->prog.binary) return c->fail(ZKB_E_ARG, "set an odd field first");
    if (!c->has_gpu) return c->fail(ZKB_E_CUDA, "no CUDA device in this context (there is no CPU fallback)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int N = c->prog.nlimb;
    const FieldParams fp = c->prog.fp;
    uint32_t* sink = nullptr;
    CUDA_TRY(c, cudaMalloc((void**)&sink, 4));
    const unsigned grid = (unsigned)c->sm_count * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_TRY(c, cudaEventRecord(c->ev[0], c->stream));
        DISPATCH_N(N, (k_field_throughput<N><<<grid, 256, 0, c->stream>>>(op, iters, sink, fp)));
        CUDA_TRY(c, cudaEventRecord(c->ev[1], c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
        if (rep > 0) best = std::min(best, ms);
    }
    cudaFree(sink);
    *ops_per_second = (double)grid * 256.0 * iters / (best * 1e-3);
    return ZKB_OK;
}
