// placeholder until evaluator.cpp lands
#include "context.h"
struct zkb_evaluator { zkb_ctx* c; std::string err; };
extern "C" zkb_evaluator* zkb_evaluator_create(zkb_ctx* b) { return new zkb_evaluator{b, ""}; }
extern "C" void zkb_evaluator_destroy(zkb_evaluator* e) { delete e; }
extern "C" int zkb_evaluator_ingest_message(zkb_evaluator*, const uint8_t*, size_t) { return ZKB_E_UNSUPPORTED; }
extern "C" int zkb_evaluator_ingest_buffer(zkb_evaluator*, const uint8_t*, size_t) { return ZKB_E_UNSUPPORTED; }
extern "C" int zkb_evaluator_ingest_paths(zkb_evaluator*, const char* const*, size_t) { return ZKB_E_UNSUPPORTED; }
extern "C" int zkb_evaluator_get_violations(zkb_evaluator*, size_t*) { return ZKB_E_UNSUPPORTED; }
extern "C" const char* zkb_evaluator_violation(zkb_evaluator*, size_t) { return nullptr; }
extern "C" int zkb_evaluator_get_wire(zkb_evaluator*, uint64_t, uint8_t*, size_t, size_t*) { return ZKB_E_UNSUPPORTED; }
extern "C" const char* zkb_evaluator_last_error(zkb_evaluator* e) { return e->err.c_str(); }
