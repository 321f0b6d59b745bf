// Device helpers shared by kernels.cu and r1cs.cu: wire-store element access (limb-chunk-major,
// witness-minor layout, see kernels.cuh) and AssertZero failure reporting.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

#include "field.cuh"
#include "field_ptx.cuh"

namespace zkb {

// ---------------------------------------------------------------------------------------------
// wire-store element access
// ---------------------------------------------------------------------------------------------
template <int N>
struct Elem {
    static constexpr int CW = N < 4 ? N : 4;  // limbs per chunk
    static constexpr int NC = N / CW;         // chunks per element
};

template <int CW>
struct Vec;
template <>
struct Vec<1> {
    using T = uint32_t;
};
template <>
struct Vec<2> {
    using T = uint2;
};
template <>
struct Vec<4> {
    using T = uint4;
};

__device__ __forceinline__ void unpack(uint32_t v, uint32_t* o) { o[0] = v; }
__device__ __forceinline__ void unpack(uint2 v, uint32_t* o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void unpack(uint4 v, uint32_t* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void pack(uint32_t& v, const uint32_t* o) { v = o[0]; }
__device__ __forceinline__ void pack(uint2& v, const uint32_t* o) { v = make_uint2(o[0], o[1]); }
__device__ __forceinline__ void pack(uint4& v, const uint32_t* o) { v = make_uint4(o[0], o[1], o[2], o[3]); }

// streaming (read-once) loads: operands are written by an earlier launch and never by this one
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) { return __ldg(p); }
__device__ __forceinline__ uint2 ld_stream(const uint2* p) { return __ldg(p); }
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldg(p); }

template <int N>
__device__ __forceinline__ void load_elem(uint32_t* r, const uint32_t* store, uint32_t slot, uint32_t lane, uint32_t log2_wt) {
    using V = typename Vec<Elem<N>::CW>::T;
    const V* base = reinterpret_cast<const V*>(store);
#pragma unroll
    for (int c = 0; c < Elem<N>::NC; c++) {
        size_t idx = (((size_t)slot * Elem<N>::NC + c) << log2_wt) + lane;
        unpack(ld_stream(base + idx), r + c * Elem<N>::CW);
    }
}

// L2-coherent variant (ld.global.cg): for kernels that read values written earlier in the SAME launch by
// other SMs (the cooperative all-levels kernel), where a stale L1 line would be wrong
__device__ __forceinline__ uint32_t ld_coherent(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ uint2 ld_coherent(const uint2* p) { return __ldcg(p); }
__device__ __forceinline__ uint4 ld_coherent(const uint4* p) { return __ldcg(p); }
template <int N>
__device__ __forceinline__ void load_elem_coherent(uint32_t* r, const uint32_t* store, uint32_t slot, uint32_t lane, uint32_t log2_wt) {
    using V = typename Vec<Elem<N>::CW>::T;
    const V* base = reinterpret_cast<const V*>(store);
#pragma unroll
    for (int c = 0; c < Elem<N>::NC; c++) {
        size_t idx = (((size_t)slot * Elem<N>::NC + c) << log2_wt) + lane;
        unpack(ld_coherent(base + idx), r + c * Elem<N>::CW);
    }
}

template <int N>
__device__ __forceinline__ void store_elem(uint32_t* store, uint32_t slot, uint32_t lane, uint32_t log2_wt, const uint32_t* r) {
    using V = typename Vec<Elem<N>::CW>::T;
    V* base = reinterpret_cast<V*>(store);
#pragma unroll
    for (int c = 0; c < Elem<N>::NC; c++) {
        size_t idx = (((size_t)slot * Elem<N>::NC + c) << log2_wt) + lane;
        V v;
        pack(v, r + c * Elem<N>::CW);
        base[idx] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// A raw little-endian value of `stride` bytes (any length: `Value`s are arbitrary byte strings,
// rust/src/structs/value.rs:11) -> its residue in Montgomery form.  Wider than one element: Horner over
// N-limb chunks, M <- M*R + chunk (in Montgomery form: mont_mul(M, R^2) + mont_mul(chunk, R^2)).
// ge_p: the raw integer is >= p, i.e. the reference would hold it unreduced (evaluator.rs:862-864).
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void load_raw_value(uint32_t* out, bool& ge_p, const uint8_t* src, uint32_t stride, const FieldParams& fp) {
    uint32_t v[N];
    if (stride == 4 * N && ((uintptr_t)src & 3) == 0) {
        if (N >= 4 && ((uintptr_t)src & 15) == 0) {  // whole elements as 16-byte vectors
            const uint4* s128 = reinterpret_cast<const uint4*>(src);
#pragma unroll
            for (int k = 0; k < N / 4; k++) unpack(__ldg(s128 + k), v + 4 * k);
        } else {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
            for (int k = 0; k < N; k++) v[k] = s32[k];
        }
        uint32_t borrow = 0;
#pragma unroll
        for (int k = 0; k < N; k++) {
            uint64_t t = (uint64_t)v[k] - fp.p[k] - borrow;
            borrow = (uint32_t)(t >> 63);
        }
        ge_p = borrow == 0;
        fe_mont_mul<N>(out, v, fp.r2, fp.p, fp.n0inv);  // also reduces v in [p, 2^(32N)) mod p
        return;
    }
    const uint32_t n_chunks = (stride + 4 * N - 1) / (4 * N);
    uint32_t acc[N];
#pragma unroll
    for (int k = 0; k < N; k++) acc[k] = 0;
    bool high_nonzero = false;
    for (uint32_t c = n_chunks; c-- > 0;) {
#pragma unroll
        for (int k = 0; k < N; k++) v[k] = 0;
        const uint32_t b0 = c * 4 * N;
        for (uint32_t b = 0; b < 4 * N && b0 + b < stride; b++) v[b >> 2] |= (uint32_t)src[b0 + b] << (8 * (b & 3));
        uint32_t m[N], t[N];
        fe_mont_mul<N>(m, v, fp.r2, fp.p, fp.n0inv);
        if (c + 1 < n_chunks) {
            fe_mont_mul<N>(t, acc, fp.r2, fp.p, fp.n0inv);
            fe_add<N>(acc, t, m, fp.p);
        } else {
#pragma unroll
            for (int k = 0; k < N; k++) acc[k] = m[k];
        }
        if (c > 0) {
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < N; k++) any |= v[k];
            high_nonzero = high_nonzero || any != 0;
        } else {
            uint32_t borrow = 0;
#pragma unroll
            for (int k = 0; k < N; k++) {
                uint64_t t2 = (uint64_t)v[k] - fp.p[k] - borrow;
                borrow = (uint32_t)(t2 >> 63);
            }
            ge_p = high_nonzero || borrow == 0;
        }
    }
#pragma unroll
    for (int k = 0; k < N; k++) out[k] = acc[k];
}

// ---------------------------------------------------------------------------------------------
// AssertZero reporting: first failing assertion (program order) per witness.
// Failures are found with one warp ballot; only failing lanes touch memory.  When all lanes of a
// warp belong to the same witness (single-witness, gate-parallel tiles) the warp first reduces its
// minimum with redux.sync and issues a single atomicMin.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void report_fail(bool fail, uint32_t seq, uint32_t* first_fail, uint32_t widx, bool single_witness) {
    unsigned act = __activemask();
    unsigned fm = __ballot_sync(act, fail);
    if (fm == 0) return;
    if (single_witness) {
        uint32_t m = __reduce_min_sync(act, fail ? seq : 0xFFFFFFFFu);
        if ((threadIdx.x & 31) == (__ffs(act) - 1)) atomicMin(first_fail + widx, m);
    } else if (fail) {
        atomicMin(first_fail + widx, seq);
    }
}

// Function attributes (cudaFuncSetAttribute) are per device: true the first time a launch site sees the current device.
// One host thread per device may be launching at the same time (zkb_evaluate_sharded), hence the atomic mask.
inline bool first_on_device(std::atomic<uint64_t>& seen) {
    int d = 0;
    cudaGetDevice(&d);
    if (d < 0 || d >= 64) return true;
    const uint64_t bit = 1ull << d;
    return (seen.fetch_or(bit) & bit) == 0;
}

}  // namespace zkb
