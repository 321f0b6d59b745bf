// Latency of reading data another SM wrote just before a grid barrier, by load flavour (the critical path of one wavefront of
// the all-levels kernel): every thread writes slot[perm(tid)], grid.sync, reads slot[perm2(tid)] with a given instruction and
// times it with clock64.  Also the same read of data written by an EARLIER launch (cold = no writer in this launch).
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

template <int MODE>
__device__ __forceinline__ uint2 ld(const uint2* p) {
    uint2 v;
    if (MODE == 0) asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else if (MODE == 1) asm volatile("ld.global.ca.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else if (MODE == 2) asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else if (MODE == 3) asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    else asm volatile("ld.global.cv.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

template <int MODE>
__global__ void k(uint2* buf, uint32_t n, int rounds, int write, long long* out, uint32_t* sink) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, T = gridDim.x * blockDim.x;
    long long lat = 0, bar = 0, wr = 0;
    uint32_t acc = 0;
    for (int r = 0; r < rounds; r++) {
        const uint32_t w = (tid * 2654435761u + r * 40503u) % n;
        long long t0 = clock64();
        if (write) buf[w] = make_uint2(tid, r);
        long long t1 = clock64();
        grid.sync();
        long long t2 = clock64();
        const uint32_t rd = ((tid + T / 2 + 12345u) * 2654435761u + r * 40503u) % n;  // a slot some far-away thread wrote this round
        uint2 v = ld<MODE>(buf + rd);
        acc += v.x + v.y;
        asm volatile("" ::"r"(acc) : "memory");
        long long t3 = clock64();
        wr += t1 - t0; bar += t2 - t1; lat += t3 - t2;
        grid.sync();
    }
    if (acc == 0x1234567u) *sink = acc;
    if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) {
        out[(blockIdx.x != 0) * 3 + 0] = wr / rounds;
        out[(blockIdx.x != 0) * 3 + 1] = bar / rounds;
        out[(blockIdx.x != 0) * 3 + 2] = lat / rounds;
    }
}

template <int MODE>
void run(const char* name, uint2* buf, uint32_t n, int blocks, int write, long long* d_out, uint32_t* sink) {
    int rounds = 50;
    void* args[] = {&buf, &n, &rounds, &write, &d_out, &sink};
    for (int i = 0; i < 2; i++) cudaLaunchCooperativeKernel((void*)k<MODE>, dim3(blocks), dim3(256), args, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[6];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-18s blocks %4d write %d: cta0 store-issue %4lld barrier %5lld load %5lld | last cta load %5lld   (%s)\n", name, blocks, write, h[0], h[1], h[2], h[5],
           cudaGetErrorString(e));
}

int main(int argc, char** argv) {
    uint32_t n = argc > 1 ? atoi(argv[1]) : (1u << 20);  // 8 MB of slots
    uint2* buf; long long* d_out; uint32_t* sink;
    cudaMalloc(&buf, (size_t)n * 8); cudaMemset(buf, 0, (size_t)n * 8);
    cudaMalloc(&d_out, 64); cudaMalloc(&sink, 4);
    for (int blocks : {148, 444}) {
        for (int write : {1, 0}) {
            run<0>("ld.global.cg", buf, n, blocks, write, d_out, sink);
            run<1>("ld.global.ca", buf, n, blocks, write, d_out, sink);
            run<2>("ld.relaxed.gpu", buf, n, blocks, write, d_out, sink);
            run<3>("ld.volatile", buf, n, blocks, write, d_out, sink);
            run<4>("ld.global.cv", buf, n, blocks, write, d_out, sink);
        }
    }
    return 0;
}
