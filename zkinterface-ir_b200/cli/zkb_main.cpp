// zkb — verbs of zki_sieve (rust/src/cli.rs) on the GPU backend.
//   zkb evaluate [--device N] <workspace dir | *.sieve ... | ->       cli.rs:130, 315-320, 557-571
//   zkb flatten --out <dir | -> <workspace dir | *.sieve ... | ->     cli.rs:442-472 (host only)
//   zkb validate <workspace dir | *.sieve ... | ->                    cli.rs:297-313 (host only)
// `evaluate` / `validate` print exactly what the reference prints on stderr ("The statement is TRUE!" /
// "The statement is COMPLIANT with the specification!" / the violation lists) and exit non-zero with
// "Found N violations." when there are any.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "zkb.h"

static int usage(const char* argv0) {
    fprintf(stderr,
            "usage: %s evaluate [--device N] <paths...>\n"
            "       %s flatten --out <dir|-> <paths...>\n"
            "       %s validate <paths...>\n",
            argv0, argv0, argv0);
    return 2;
}

// print_violations, cli.rs:557-571
static int print_violations(size_t n, const std::vector<std::string>& v, const char* what_it_is_supposed_to_be) {
    fprintf(stderr, "\n");
    if (n > 0) {
        fprintf(stderr, "The statement is NOT %s!\n", what_it_is_supposed_to_be);
        fprintf(stderr, "Violations:\n");
        for (const auto& s : v) fprintf(stderr, "- %s\n", s.c_str());
        fprintf(stderr, "\nError: Found %zu violations.\n", n);
        return 1;
    }
    fprintf(stderr, "The statement is %s!\n", what_it_is_supposed_to_be);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) return usage(argv[0]);
    const std::string verb = argv[1];
    int device = 0;
    const char* out = nullptr;
    std::vector<const char*> paths;
    for (int i = 2; i < argc; i++) {
        if (strcmp(argv[i], "--device") == 0 && i + 1 < argc) device = atoi(argv[++i]);
        else if (strcmp(argv[i], "--out") == 0 && i + 1 < argc) out = argv[++i];
        else paths.push_back(argv[i]);
    }
    if (verb == "evaluate") {
        zkb_ctx* ctx = zkb_create(device);
        if (zkb_last_error(ctx)[0]) {
            fprintf(stderr, "Error: %s\n", zkb_last_error(ctx));
            return 1;
        }
        zkb_evaluator* ev = zkb_evaluator_create(ctx);
        int rc = zkb_evaluator_ingest_paths(ev, paths.data(), paths.size());
        size_t n = 0;
        if (rc == ZKB_OK) rc = zkb_evaluator_get_violations(ev, &n);
        if (rc != ZKB_OK) {
            fprintf(stderr, "Error: %s\n", zkb_evaluator_last_error(ev));
            return 1;
        }
        std::vector<std::string> v;
        for (size_t i = 0; i < n; i++) v.push_back(zkb_evaluator_violation(ev, i));
        int status = print_violations(n, v, "TRUE");
        zkb_evaluator_destroy(ev);
        zkb_destroy(ctx);
        return status;
    }
    if (verb == "flatten") {
        if (!out) return usage(argv[0]);
        zkb_ctx* ctx = zkb_create(-1);
        zkb_evaluator* ev = zkb_evaluator_create(ctx);
        int rc = zkb_evaluator_set_flatten(ev, 1);
        if (rc == ZKB_OK) rc = zkb_evaluator_ingest_paths(ev, paths.data(), paths.size());
        if (rc == ZKB_OK) {
            if (strcmp(out, "-") == 0) {  // MemorySink, then instance / witness / relation messages to stdout
                const uint8_t* p[3];
                size_t n[3];
                rc = zkb_evaluator_flatten(ev, &p[0], &n[0], &p[1], &n[1], &p[2], &n[2]);
                for (int i = 0; i < 3 && rc == ZKB_OK; i++)
                    if (n[i]) fwrite(p[i], 1, n[i], stdout);
            } else {
                rc = zkb_evaluator_flatten_to_dir(ev, out);
            }
        }
        if (rc != ZKB_OK) {
            fprintf(stderr, "Error: %s\n", zkb_evaluator_last_error(ev));
            return 1;
        }
        zkb_evaluator_destroy(ev);
        zkb_destroy(ctx);
        return 0;
    }
    if (verb == "validate") {
        zkb_validator* val = zkb_validator_create(1);  // main_validate validates as prover (cli.rs:301)
        int rc = zkb_validator_ingest_paths(val, paths.data(), paths.size());
        size_t n = 0;
        if (rc == ZKB_OK) rc = zkb_validator_get_violations(val, &n);
        if (rc != ZKB_OK) {
            fprintf(stderr, "Error: %s\n", zkb_validator_last_error(val));
            return 1;
        }
        std::vector<std::string> v;
        for (size_t i = 0; i < n; i++) v.push_back(zkb_validator_violation(val, i));
        int status = print_violations(n, v, "COMPLIANT with the specification");
        zkb_validator_destroy(val);
        return status;
    }
    return usage(argv[0]);
}
