// Device helpers shared by kernels.cu and r1cs.cu: wire-store element access (limb-chunk-major,
// witness-minor layout, see kernels.cuh) and AssertZero failure reporting.
#pragma once
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

}  // namespace zkb
