/* plaintext_flat.c — CPU restatement of the reference's gate loop for FLAT relations.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py).  Never linked into
 * or called by the product (zkinterface-ir_b200/).
 *
 * Restates, for relations made of simple gates only, what `zki_sieve evaluate` does:
 *   Evaluator::ingest_gate simple arms   rust/src/consumers/evaluator.rs:344-439
 *   get / set / remove on the scope map  rust/src/consumers/evaluator.rs:775-797
 *   PlaintextBackend arithmetic          rust/src/consumers/evaluator.rs:862-946
 * and keeps the reference's STRUCTURE on purpose, so that timing it is a fair stand-in
 * for the Rust binary (which cannot be built here: no cargo/rustc):
 *   - the wire store is a hash map keyed by the u64 wire id (HashMap<WireId, BigUint>),
 *   - every value is a heap-allocated little-endian digit vector (BigUint),
 *   - add/mul compute the full-width integer, then a generic long division takes `% m`
 *     (num-bigint 0.3.0 `%` = Knuth algorithm D; restated below),
 *   - constants are re-parsed from bytes on every execution (evaluator.rs:388,395),
 *   - evaluation stops at the first failing assertion (evaluator.rs:214-221, 357-362).
 * Field elements are NOT reduced on input (evaluator.rs:862-864, 896-898).
 *
 * Pinned against oracle/evaluator.py (itself pinned on the reference's golden vectors)
 * by tests/test_oracle_flat.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint8_t op;
    uint8_t pad[3];
    uint32_t out, a, b;
} flat_gate; /* same layout as zkb_gate (include/zkb.h); opcodes = DirectiveSet 1..13 */

enum { G_CONSTANT = 1, G_ASSERT_ZERO, G_COPY, G_ADD, G_MUL, G_ADD_CONSTANT, G_MUL_CONSTANT, G_AND, G_XOR, G_NOT,
       G_INSTANCE, G_WITNESS, G_FREE };

/* ---------------------------------------------------------------- big unsigned */
typedef struct {
    uint32_t* d; /* little-endian digits, normalised (no leading zeros) */
    int n;
} bn;

static bn bn_alloc(int n) {
    bn r;
    r.d = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n > 0 ? n : 1));
    r.n = n;
    return r;
}
static void bn_free(bn* a) {
    free(a->d);
    a->d = NULL;
    a->n = 0;
}
static void bn_norm(bn* a) {
    while (a->n > 0 && a->d[a->n - 1] == 0) a->n--;
}
static bn bn_from_bytes_le(const uint8_t* b, size_t len) { /* BigUint::from_bytes_le */
    bn r = bn_alloc((int)((len + 3) / 4));
    memset(r.d, 0, sizeof(uint32_t) * (size_t)(r.n > 0 ? r.n : 1));
    for (size_t i = 0; i < len; i++) r.d[i / 4] |= (uint32_t)b[i] << (8 * (i % 4));
    bn_norm(&r);
    return r;
}
static bn bn_clone(const bn* a) {
    bn r = bn_alloc(a->n);
    memcpy(r.d, a->d, sizeof(uint32_t) * (size_t)a->n);
    return r;
}
static int bn_cmp(const bn* a, const bn* b) {
    if (a->n != b->n) return a->n < b->n ? -1 : 1;
    for (int i = a->n - 1; i >= 0; i--)
        if (a->d[i] != b->d[i]) return a->d[i] < b->d[i] ? -1 : 1;
    return 0;
}
static bn bn_add(const bn* a, const bn* b) {
    if (a->n < b->n) { const bn* t = a; a = b; b = t; }
    bn r = bn_alloc(a->n + 1);
    uint64_t c = 0;
    for (int i = 0; i < a->n; i++) {
        c += (uint64_t)a->d[i] + (i < b->n ? b->d[i] : 0);
        r.d[i] = (uint32_t)c;
        c >>= 32;
    }
    r.d[a->n] = (uint32_t)c;
    bn_norm(&r);
    return r;
}
static bn bn_mul(const bn* a, const bn* b) {
    if (a->n == 0 || b->n == 0) return bn_alloc(0);
    bn r = bn_alloc(a->n + b->n);
    memset(r.d, 0, sizeof(uint32_t) * (size_t)r.n);
    for (int i = 0; i < a->n; i++) {
        uint64_t c = 0;
        for (int j = 0; j < b->n; j++) {
            c += (uint64_t)a->d[i] * b->d[j] + r.d[i + j];
            r.d[i + j] = (uint32_t)c;
            c >>= 32;
        }
        r.d[i + b->n] = (uint32_t)c;
    }
    bn_norm(&r);
    return r;
}
/* a % m, m != 0: Knuth TAOCP vol.2 4.3.1 algorithm D (what num-bigint's div_rem implements) */
static bn bn_mod(const bn* a, const bn* m) {
    if (bn_cmp(a, m) < 0) return bn_clone(a);
    if (m->n == 1) {
        uint64_t rem = 0;
        for (int i = a->n - 1; i >= 0; i--) rem = ((rem << 32) | a->d[i]) % m->d[0];
        bn r = bn_alloc(1);
        r.d[0] = (uint32_t)rem;
        bn_norm(&r);
        return r;
    }
    int s = __builtin_clz(m->d[m->n - 1]);
    int n = m->n, mm = a->n - m->n;
    uint32_t* v = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* u = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(a->n + 1));
    for (int i = n - 1; i > 0; i--) v[i] = s ? (m->d[i] << s) | (m->d[i - 1] >> (32 - s)) : m->d[i];
    v[0] = m->d[0] << s;
    u[a->n] = s ? a->d[a->n - 1] >> (32 - s) : 0;
    for (int i = a->n - 1; i > 0; i--) u[i] = s ? (a->d[i] << s) | (a->d[i - 1] >> (32 - s)) : a->d[i];
    u[0] = a->d[0] << s;
    for (int j = mm; j >= 0; j--) {
        uint64_t num = ((uint64_t)u[j + n] << 32) | u[j + n - 1];
        uint64_t qhat = num / v[n - 1], rhat = num % v[n - 1];
        while (qhat >= (1ull << 32) || qhat * v[n - 2] > ((rhat << 32) | u[j + n - 2])) {
            qhat--;
            rhat += v[n - 1];
            if (rhat >= (1ull << 32)) break;
        }
        int64_t borrow = 0;
        uint64_t carry = 0;
        for (int i = 0; i < n; i++) {
            uint64_t p = qhat * v[i] + carry;
            carry = p >> 32;
            int64_t t = (int64_t)u[i + j] - borrow - (int64_t)(p & 0xFFFFFFFFull);
            u[i + j] = (uint32_t)t;
            borrow = t < 0 ? 1 : 0;
        }
        int64_t t = (int64_t)u[j + n] - borrow - (int64_t)carry;
        u[j + n] = (uint32_t)t;
        if (t < 0) { /* add back */
            uint64_t c = 0;
            for (int i = 0; i < n; i++) {
                c += (uint64_t)u[i + j] + v[i];
                u[i + j] = (uint32_t)c;
                c >>= 32;
            }
            u[j + n] += (uint32_t)c;
        }
    }
    bn r = bn_alloc(n);
    for (int i = 0; i < n - 1; i++) r.d[i] = s ? (u[i] >> s) | (u[i + 1] << (32 - s)) : u[i];
    r.d[n - 1] = u[n - 1] >> s;
    bn_norm(&r);
    free(u);
    free(v);
    return r;
}
static bn bn_bitop(const bn* a, const bn* b, int is_xor) {
    int n = a->n > b->n ? a->n : b->n;
    bn r = bn_alloc(n);
    for (int i = 0; i < n; i++) {
        uint32_t x = i < a->n ? a->d[i] : 0, y = i < b->n ? b->d[i] : 0;
        r.d[i] = is_xor ? (x ^ y) : (x & y);
    }
    bn_norm(&r);
    return r;
}

/* ---------------------------------------------------------------- scope: HashMap<WireId, BigUint> */
typedef struct {
    uint64_t* key;
    bn* val;
    uint8_t* state; /* 0 empty, 1 full, 2 tombstone */
    size_t cap, used, filled;
} scope_t;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
static void scope_init(scope_t* s, size_t cap) {
    s->cap = cap;
    s->used = s->filled = 0;
    s->key = (uint64_t*)malloc(sizeof(uint64_t) * cap);
    s->val = (bn*)malloc(sizeof(bn) * cap);
    s->state = (uint8_t*)calloc(cap, 1);
}
static void scope_destroy(scope_t* s) {
    for (size_t i = 0; i < s->cap; i++)
        if (s->state[i] == 1) bn_free(&s->val[i]);
    free(s->key); free(s->val); free(s->state);
}
static bn* scope_get(scope_t* s, uint64_t k) {
    size_t i = mix64(k) & (s->cap - 1);
    while (s->state[i]) {
        if (s->state[i] == 1 && s->key[i] == k) return &s->val[i];
        i = (i + 1) & (s->cap - 1);
    }
    return NULL;
}
static void scope_grow(scope_t* s);
/* returns 0 if the key was already present (value replaced, like HashMap::insert) */
static int scope_set(scope_t* s, uint64_t k, bn v) {
    bn* e = scope_get(s, k);
    if (e) { bn_free(e); *e = v; return 0; }
    if ((s->filled + 1) * 10 > s->cap * 7) scope_grow(s);
    size_t i = mix64(k) & (s->cap - 1);
    while (s->state[i] == 1) i = (i + 1) & (s->cap - 1);
    if (s->state[i] == 0) s->filled++;
    s->state[i] = 1; s->key[i] = k; s->val[i] = v; s->used++;
    return 1;
}
static void scope_grow(scope_t* s) {
    scope_t n;
    scope_init(&n, s->cap * 2);
    for (size_t i = 0; i < s->cap; i++)
        if (s->state[i] == 1) scope_set(&n, s->key[i], s->val[i]);
    free(s->key); free(s->val); free(s->state);
    *s = n;
}
static int scope_remove(scope_t* s, uint64_t k) {
    size_t i = mix64(k) & (s->cap - 1);
    while (s->state[i]) {
        if (s->state[i] == 1 && s->key[i] == k) { bn_free(&s->val[i]); s->state[i] = 2; s->used--; return 1; }
        i = (i + 1) & (s->cap - 1);
    }
    return 0;
}

/* ---------------------------------------------------------------- evaluation of one statement */
/* status codes of flat_eval */
enum { EV_TRUE = 0, EV_ASSERT_FAILED = 1, EV_NO_VALUE = 2, EV_ALREADY_SET = 3, EV_NOT_ENOUGH_INSTANCE = 4,
       EV_MISSING_WITNESS_PANIC = 5, EV_BAD_GATE = 6 };

typedef struct {
    int32_t status;
    uint32_t pad;
    uint64_t fail_assert_seq; /* EV_ASSERT_FAILED: index (program order) of the failing AssertZero */
    uint64_t fail_wire;       /* wire id named in the reference's error text */
    uint64_t gates_done;      /* gates ingested before stopping */
} flat_result;

/* dump (optional): value of wire id w (if still live at the end) at dump + w*dump_stride */
static void flat_eval_one(const flat_gate* g, uint64_t n_gates, const uint8_t* cpool, size_t cstride, const uint8_t* modulus,
                          size_t mod_len, const uint8_t* inst, uint64_t n_inst, const uint8_t* wit, uint64_t n_wit,
                          size_t vstride, flat_result* res, uint8_t* dump, size_t dump_stride, uint64_t dump_wires) {
    bn m = bn_from_bytes_le(modulus, mod_len);
    scope_t sc;
    scope_init(&sc, 1024);
    uint64_t ipos = 0, wpos = 0, aseq = 0;
    memset(res, 0, sizeof(*res));
    res->fail_assert_seq = UINT64_MAX;
    uint64_t i = 0;
#define FAIL(code, wire) do { res->status = (code); res->fail_wire = (wire); goto done; } while (0)
    for (; i < n_gates; i++) {
        const flat_gate* q = &g[i];
        switch (q->op) {
            case G_CONSTANT: { /* :345-348 */
                bn v = bn_from_bytes_le(cpool + (size_t)q->b * cstride, cstride);
                if (!scope_set(&sc, q->out, v)) FAIL(EV_ALREADY_SET, q->out);
            } break;
            case G_ASSERT_ZERO: { /* :350-364, unweighted: copy then is_zero */
                bn* a = scope_get(&sc, q->a);
                if (!a) FAIL(EV_NO_VALUE, q->a);
                bn c = bn_clone(a);
                int zero = c.n == 0;
                bn_free(&c);
                if (!zero) { res->fail_assert_seq = aseq; FAIL(EV_ASSERT_FAILED, q->a); }
                aseq++;
            } break;
            case G_COPY: { /* :366-370 */
                bn* a = scope_get(&sc, q->a);
                if (!a) FAIL(EV_NO_VALUE, q->a);
                if (!scope_set(&sc, q->out, bn_clone(a))) FAIL(EV_ALREADY_SET, q->out);
            } break;
            case G_ADD: case G_MUL: case G_AND: case G_XOR: { /* :372-384, 400-412 ; backend :908-930 */
                bn* a = scope_get(&sc, q->a);
                if (!a) FAIL(EV_NO_VALUE, q->a);
                bn* b = scope_get(&sc, q->b);
                if (!b) FAIL(EV_NO_VALUE, q->b);
                bn t = q->op == G_ADD ? bn_add(a, b) : q->op == G_MUL ? bn_mul(a, b) : bn_bitop(a, b, q->op == G_XOR);
                bn r = bn_mod(&t, &m);
                bn_free(&t);
                if (!scope_set(&sc, q->out, r)) FAIL(EV_ALREADY_SET, q->out);
            } break;
            case G_ADD_CONSTANT: case G_MUL_CONSTANT: { /* :386-398 ; backend :916-922 */
                bn* a = scope_get(&sc, q->a);
                if (!a) FAIL(EV_NO_VALUE, q->a);
                bn c = bn_from_bytes_le(cpool + (size_t)q->b * cstride, cstride);
                bn t = q->op == G_ADD_CONSTANT ? bn_add(a, &c) : bn_mul(a, &c);
                bn r = bn_mod(&t, &m);
                bn_free(&t); bn_free(&c);
                if (!scope_set(&sc, q->out, r)) FAIL(EV_ALREADY_SET, q->out);
            } break;
            case G_NOT: { /* :414-418 ; backend :932-938 */
                bn* a = scope_get(&sc, q->a);
                if (!a) FAIL(EV_NO_VALUE, q->a);
                bn r = bn_alloc(1);
                r.d[0] = 1;
                r.n = a->n == 0 ? 1 : 0;
                if (!scope_set(&sc, q->out, r)) FAIL(EV_ALREADY_SET, q->out);
            } break;
            case G_INSTANCE: { /* :420-427 */
                if (ipos >= n_inst) FAIL(EV_NOT_ENOUGH_INSTANCE, q->out);
                bn v = bn_from_bytes_le(inst + (size_t)ipos++ * vstride, vstride);
                if (!scope_set(&sc, q->out, v)) FAIL(EV_ALREADY_SET, q->out);
            } break;
            case G_WITNESS: { /* :429-432, backend :944-946 panics on None */
                if (wpos >= n_wit) FAIL(EV_MISSING_WITNESS_PANIC, q->out);
                bn v = bn_from_bytes_le(wit + (size_t)wpos++ * vstride, vstride);
                if (!scope_set(&sc, q->out, v)) FAIL(EV_ALREADY_SET, q->out);
            } break;
            case G_FREE: /* :434-439 */
                for (uint64_t w = q->a; w <= (uint64_t)q->b; w++)
                    if (!scope_remove(&sc, w)) FAIL(EV_NO_VALUE, w);
                break;
            default: FAIL(EV_BAD_GATE, i);
        }
    }
done:
    res->gates_done = i;
    if (dump) {
        memset(dump, 0, dump_stride * dump_wires);
        for (uint64_t w = 0; w < dump_wires; w++) {
            bn* a = scope_get(&sc, w);
            if (!a) { memset(dump + w * dump_stride, 0xFF, dump_stride); continue; } /* not live */
            for (int k = 0; k < a->n && (size_t)k * 4 < dump_stride; k++)
                for (int b = 0; b < 4 && (size_t)k * 4 + b < dump_stride; b++)
                    dump[w * dump_stride + (size_t)k * 4 + b] = (uint8_t)(a->d[k] >> (8 * b));
        }
    }
    scope_destroy(&sc);
    bn_free(&m);
}

typedef struct {
    const flat_gate* g; uint64_t n_gates; const uint8_t* cpool; size_t cstride; const uint8_t* modulus; size_t mod_len;
    const uint8_t* inst; uint64_t inst_set_stride; uint64_t n_inst; const uint8_t* wit; uint64_t wit_set_stride;
    uint64_t n_wit; size_t vstride; flat_result* res; uint32_t first, last;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    for (uint32_t k = j->first; k < j->last; k++)
        flat_eval_one(j->g, j->n_gates, j->cpool, j->cstride, j->modulus, j->mod_len, j->inst + (size_t)k * j->inst_set_stride,
                      j->n_inst, j->wit + (size_t)k * j->wit_set_stride, j->n_wit, j->vstride, &j->res[k], NULL, 0, 0);
    return NULL;
}

/* evaluate n_batch independent (instance, witness) pairs against one flat relation, on n_threads threads */
int flat_eval_batch(const flat_gate* g, uint64_t n_gates, const uint8_t* cpool, size_t cstride, const uint8_t* modulus,
                    size_t mod_len, const uint8_t* inst, uint64_t inst_set_stride, uint64_t n_inst, const uint8_t* wit,
                    uint64_t wit_set_stride, uint64_t n_wit, size_t vstride, uint32_t n_batch, int n_threads, flat_result* res) {
    if (n_threads < 1) n_threads = 1;
    if ((uint32_t)n_threads > n_batch) n_threads = (int)n_batch;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    job_t* jobs = (job_t*)malloc(sizeof(job_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) {
        job_t j = {g, n_gates, cpool, cstride, modulus, mod_len, inst, inst_set_stride, n_inst, wit, wit_set_stride, n_wit,
                   vstride, res, (uint32_t)((uint64_t)n_batch * t / n_threads), (uint32_t)((uint64_t)n_batch * (t + 1) / n_threads)};
        jobs[t] = j;
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}

/* single statement, with a dump of the first dump_wires wire ids (0xFF.. = not live) */
int flat_eval_dump(const flat_gate* g, uint64_t n_gates, const uint8_t* cpool, size_t cstride, const uint8_t* modulus,
                   size_t mod_len, const uint8_t* inst, uint64_t n_inst, const uint8_t* wit, uint64_t n_wit, size_t vstride,
                   flat_result* res, uint8_t* dump, size_t dump_stride, uint64_t dump_wires) {
    flat_eval_one(g, n_gates, cpool, cstride, modulus, mod_len, inst, n_inst, wit, n_wit, vstride, res, dump, dump_stride, dump_wires);
    return 0;
}

/* the big-integer primitives, exported so the tests can pin them against Python ints */
int flat_bn_mulmod(const uint8_t* a, size_t al, const uint8_t* b, size_t bl, const uint8_t* m, size_t ml, uint8_t* out, size_t ol) {
    bn x = bn_from_bytes_le(a, al), y = bn_from_bytes_le(b, bl), mm = bn_from_bytes_le(m, ml);
    bn t = bn_mul(&x, &y);
    bn r = bn_mod(&t, &mm);
    memset(out, 0, ol);
    for (int k = 0; k < r.n; k++)
        for (int q = 0; q < 4 && (size_t)k * 4 + q < ol; q++) out[(size_t)k * 4 + q] = (uint8_t)(r.d[k] >> (8 * q));
    bn_free(&x); bn_free(&y); bn_free(&mm); bn_free(&t); bn_free(&r);
    return 0;
}
