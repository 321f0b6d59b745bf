// zkb — the `evaluate` verb of zki_sieve (rust/src/cli.rs:130, 315-320, 557-571) on the GPU backend.
//   zkb evaluate [--device N] <workspace dir | *.sieve ...>
// Prints exactly what the reference prints on stderr ("The statement is TRUE!" / "The statement is
// NOT TRUE!" + the violation list) and exits non-zero with "Found N violations." when there are any.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "zkb.h"

int main(int argc, char** argv) {
    if (argc < 3 || strcmp(argv[1], "evaluate") != 0) {
        fprintf(stderr, "usage: %s evaluate [--device N] <paths...>\n", argv[0]);
        return 2;
    }
    int device = 0;
    std::vector<const char*> paths;
    for (int i = 2; i < argc; i++) {
        if (strcmp(argv[i], "--device") == 0 && i + 1 < argc) device = atoi(argv[++i]);
        else paths.push_back(argv[i]);
    }
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
    fprintf(stderr, "\n");
    int status = 0;
    if (n > 0) {
        fprintf(stderr, "The statement is NOT TRUE!\n");
        fprintf(stderr, "Violations:\n");
        for (size_t i = 0; i < n; i++) fprintf(stderr, "- %s\n", zkb_evaluator_violation(ev, i));
        fprintf(stderr, "\nError: Found %zu violations.\n", n);
        status = 1;
    } else {
        fprintf(stderr, "The statement is TRUE!\n");
    }
    zkb_evaluator_destroy(ev);
    zkb_destroy(ctx);
    return status;
}
