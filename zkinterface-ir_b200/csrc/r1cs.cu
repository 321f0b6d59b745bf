// R1CS Az o Bz = Cz check (section 5 of include/zkb.h) — see r1cs kernels below.
#include "context.h"
using namespace zkb;
namespace zkb {
struct R1csDev {};
void r1cs_free(zkb_ctx* c) { delete c->r1cs; c->r1cs = nullptr; }
}
extern "C" int zkb_r1cs_load(zkb_ctx* c, const zkb_csr*, const zkb_csr*, const zkb_csr*, const uint8_t*, size_t, uint64_t, uint64_t) { return c->fail(ZKB_E_UNSUPPORTED, "r1cs: not built yet"); }
extern "C" int zkb_r1cs_check(zkb_ctx* c, const uint8_t*, uint64_t, uint32_t, uint32_t, zkb_verdict*) { return c->fail(ZKB_E_UNSUPPORTED, "r1cs: not built yet"); }
extern "C" int zkb_r1cs_upload(zkb_ctx* c, const uint8_t*, uint64_t, uint32_t, uint32_t) { return c->fail(ZKB_E_UNSUPPORTED, "r1cs: not built yet"); }
extern "C" int zkb_r1cs_run(zkb_ctx* c, zkb_verdict*) { return c->fail(ZKB_E_UNSUPPORTED, "r1cs: not built yet"); }
