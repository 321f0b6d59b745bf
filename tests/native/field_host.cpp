// Host build of zkinterface-ir_b200/csrc/field.cuh + Program::set_field for unit tests (no GPU):
// the same __host__ __device__ arithmetic the kernels run, checked against Python integers.
#include <stdint.h>
#include <string.h>

#include <string>

#include "../../zkinterface-ir_b200/csrc/field.cuh"
#include "../../zkinterface-ir_b200/csrc/program.h"

using namespace zkb;

template <int N>
static void run(int op, const FieldParams& fp, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    switch (op) {
        case 0: fe_add<N>(r, a, b, fp.p); break;
        case 1: fe_mont_mul_portable<N>(r, a, b, fp.p, fp.n0inv); break;
        case 2: {  // full modular product: to Montgomery, multiply, back
            uint32_t am[N], bm[N], pm[N], one[N];
            for (int k = 0; k < N; k++) one[k] = (k == 0);
            fe_mont_mul_portable<N>(am, a, fp.r2, fp.p, fp.n0inv);
            fe_mont_mul_portable<N>(bm, b, fp.r2, fp.p, fp.n0inv);
            fe_mont_mul_portable<N>(pm, am, bm, fp.p, fp.n0inv);
            fe_mont_mul_portable<N>(r, pm, one, fp.p, fp.n0inv);
        } break;
        case 3: fe_and_canon<N>(r, a, b); break;
        case 4: fe_xor_canon<N>(r, a, b, fp.p); break;
    }
}

extern "C" int field_host_params(const uint8_t* mod_le, size_t len, FieldParams* out) {
    Program p;
    std::string err;
    if (!p.set_field(mod_le, len, 1, err)) return -1;
    *out = p.fp;
    return p.nlimb;
}

// n operations on n (a, b) pairs of nlimb limbs each
extern "C" void field_host_batch(int op, const FieldParams* fp, const uint32_t* a, const uint32_t* b, uint32_t* r, size_t n) {
    int N = (int)fp->nlimb;
    for (size_t i = 0; i < n; i++) {
        const uint32_t *x = a + i * N, *y = b + i * N;
        uint32_t* z = r + i * N;
        if (N == 1) run<1>(op, *fp, x, y, z);
        else if (N == 2) run<2>(op, *fp, x, y, z);
        else if (N == 4) run<4>(op, *fp, x, y, z);
        else run<8>(op, *fp, x, y, z);
    }
}
