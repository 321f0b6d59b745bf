// zkb_evaluator — mirror of `Evaluator<B>` (rust/src/consumers/evaluator.rs:158-753) and of
// `Source` (rust/src/consumers/source.rs:45-193) over the recording GPU backend.
//
// The control flow of ingest_gate / ingest_subcircuit / compute_weight / exp is reproduced
// callback for callback (so the recorded program is op-for-op what the reference would ask of a
// ZKBackend), but nothing is evaluated here: assertions are resolved by the device and the
// reference's violation strings are rebuilt in get_violations from the first failing assertion.
#include <dirent.h>
#include <stdio.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <deque>
#include <unordered_map>

#include "context.h"
#include "ir.h"

using namespace zkb;

namespace {

struct Fatal {  // conditions on which the reference panics
    std::string msg;
};
struct EvalErr {  // Err(Box<dyn Error>) of the reference
    std::string msg;
};

struct FunctionDecl {  // evaluator.rs:130-136
    std::shared_ptr<std::vector<ir::Gate>> body;
    std::shared_ptr<std::vector<std::vector<uint8_t>>> consts;  // the message's constant table the body indexes
    uint64_t instance_nbr, witness_nbr, output_count, input_count;
};

using Queue = std::deque<uint32_t>;  // positions in the instance / witness value streams
using Iters = std::vector<std::pair<std::string, uint64_t>>;

std::string u64s(uint64_t v) { return std::to_string((unsigned long long)v); }

}  // namespace

struct zkb_evaluator {
    zkb_ctx* c;
    std::string err;

    // Evaluator state, evaluator.rs:158-170
    Scope values;
    std::vector<uint8_t> modulus_le;
    Queue instance_queue, witness_queue;
    std::vector<std::vector<uint8_t>> instance_values, witness_values;  // the streams themselves
    bool is_boolean = false;
    std::unordered_map<std::string, FunctionDecl> known_functions;
    bool verified_at_least_one_gate = false;
    bool has_error = false;
    std::string found_error;
    bool fatal = false;

    bool evaluated = false;
    std::vector<std::string> violations;
    std::vector<Scope*> scope_pool;

    explicit zkb_evaluator(zkb_ctx* ctx) : c(ctx) {}
    ~zkb_evaluator() {
        for (auto s : scope_pool) delete s;
    }

    Program& prog() { return c->prog; }

    // work accounting against zkb_set_limits (not in the reference, which would run until memory is exhausted)
    uint64_t steps = 0;
    void step(uint64_t n = 1) {
        steps += n;
        if (steps > c->max_steps || steps < n) throw EvalErr{"zkb: resource limit exceeded (max_steps)"};
    }

    Scope* new_scope() {
        if (scope_pool.empty()) return new Scope();
        Scope* s = scope_pool.back();
        scope_pool.pop_back();
        return s;
    }
    void release_scope(Scope* s) {
        s->clear();
        scope_pool.push_back(s);
    }

    // ---- scope helpers, evaluator.rs:775-797 -------------------------------------------------
    static uint32_t get(const Scope& sc, uint64_t id) {
        uint32_t v = sc.get(id);
        if (v == Scope::kNone) throw EvalErr{"No value given for wire_" + u64s(id)};
        return v;
    }
    static void set(Scope& sc, uint64_t id, uint32_t v) {
        if (!sc.set(id, v)) throw EvalErr{"Wire_" + u64s(id) + " already has a value in this scope."};
    }
    static void remove(Scope& sc, uint64_t id) {
        if (!sc.remove(id)) throw EvalErr{"No value given for wire_" + u64s(id)};
    }

    // ---- structs/wire.rs:179-203 ---------------------------------------------------------------
    void expand_wirelist(const ir::WireList& wl, std::vector<uint64_t>& out) {
        out.clear();
        for (const auto& e : wl) {
            if (!e.is_range) {
                out.push_back(e.first);
            } else {
                if (e.last <= e.first)
                    throw EvalErr{"In WireRange, last WireId (" + u64s(e.last) + ") must be strictly greater than first WireId (" +
                                  u64s(e.first) + ")."};
                if (e.last - e.first > (1ull << 32)) throw EvalErr{"zkb: wire range too large"};
                step(e.last - e.first);
                for (uint64_t w = e.first; w <= e.last; w++) out.push_back(w);
            }
        }
    }

    // ---- structs/iterators.rs:349-403 ----------------------------------------------------------
    static uint64_t eval_iterexpr(const ir::IterExpr& e, const Iters& known) {
        switch (e.type) {
            case 1: return e.value;
            case 2:
                for (size_t i = known.size(); i-- > 0;)
                    if (known[i].first == e.name) return known[i].second;
                // evaluate_iterexpr_list unwraps the Err: a panic in the reference (iterators.rs:400)
                throw Fatal{"Unknown iterator name " + e.name};
            case 3: return eval_iterexpr(*e.l, known) + eval_iterexpr(*e.r, known);  // release build: wrapping
            case 4: return eval_iterexpr(*e.l, known) - eval_iterexpr(*e.r, known);
            case 5: return eval_iterexpr(*e.l, known) * eval_iterexpr(*e.r, known);
            case 6:
                if (e.value == 0) throw Fatal{"attempt to divide by zero"};
                return eval_iterexpr(*e.l, known) / e.value;
        }
        throw Fatal{"Unknown Iterator Expression type"};
    }
    void eval_iterexpr_list(const ir::IterExprList& l, const Iters& known, std::vector<uint64_t>& out) {
        out.clear();
        for (const auto& el : l) {
            uint64_t a = eval_iterexpr(el.first, known);
            if (!el.is_range) {
                out.push_back(a);
            } else {
                uint64_t b = eval_iterexpr(el.last, known);
                if (b >= a && b - a > (1ull << 32)) throw EvalErr{"zkb: iterator range too large"};
                if (b >= a) step(b - a);
                for (uint64_t w = a; w <= b && b >= a; w++) {
                    out.push_back(w);
                    if (w == UINT64_MAX) break;
                }
            }
        }
    }
    static void iters_set(Iters& k, const std::string& name, uint64_t v) {  // HashMap::insert
        for (auto& kv : k)
            if (kv.first == name) {
                kv.second = v;
                return;
            }
        k.push_back({name, v});
    }
    static void iters_remove(Iters& k, const std::string& name) {
        for (size_t i = 0; i < k.size(); i++)
            if (k[i].first == name) {
                k.erase(k.begin() + i);
                return;
            }
    }

    // ---- evaluator.rs:78-126 -------------------------------------------------------------------
    uint32_t as_mul(uint32_t a, uint32_t b) { return is_boolean ? prog().and_(a, b) : prog().multiply(a, b); }
    uint32_t as_add(uint32_t a, uint32_t b) { return is_boolean ? prog().xor_(a, b) : prog().add(a, b); }
    uint32_t as_negate(uint32_t w) {
        if (is_boolean) return prog().copy(w);
        std::vector<uint8_t> m1 = prog().minus_one_le();
        return prog().mul_constant(w, m1.data(), m1.size());
    }
    uint32_t as_add_one(uint32_t w) {
        if (is_boolean) return prog().not_(w);
        uint8_t one = 1;
        return prog().add_constant(w, &one, 1);
    }
    // evaluator.rs:801-820: recursive square-and-multiply, MSB -> LSB; the recursion bottoms out in
    // copy(base) at exponent 1 and then, per lower bit, squares and (bit set) multiplies by base.
    uint32_t exp(uint32_t base, const BigU& exponent) {
        size_t nb = exponent.bits();
        uint32_t acc = prog().copy(base);
        for (size_t i = nb - 1; i-- > 0;) {
            acc = as_mul(acc, acc);
            if (exponent.bit(i)) acc = as_mul(acc, base);
        }
        return acc;
    }
    // evaluator.rs:823-839
    uint32_t compute_weight(const std::vector<uint8_t>& case_val, uint32_t condition) {
        uint32_t case_wire = prog().constant(case_val.data(), case_val.size());
        BigU exponent = BigU::from_bytes_le(modulus_le.data(), modulus_le.size());
        exponent.sub(BigU(1));
        uint32_t minus_cond = as_negate(condition);
        uint32_t base = as_add(case_wire, minus_cond);
        uint32_t base_to_exp = exp(base, exponent);
        uint32_t right = as_negate(base_to_exp);
        return as_add_one(right);
    }

    void check_arity(const std::string& name, const FunctionDecl& f, size_t n_out, size_t n_in) {  // :449-454
        if (n_out != f.output_count)
            throw EvalErr{"Wrong number of output variables in call to function " + name + " (Expected " + u64s(f.output_count) +
                          " / Got " + u64s(n_out) + ")."};
        if (n_in != f.input_count)
            throw EvalErr{"Wrong number of input variables in call to function " + name + " (Expected " + u64s(f.input_count) +
                          " / Got " + u64s(n_in) + ")."};
    }

    // ---- evaluator.rs:698-746 --------------------------------------------------------------------
    void ingest_subcircuit(const std::vector<ir::Gate>& sub, const std::vector<std::vector<uint8_t>>& consts,
                           const std::vector<uint64_t>& outs, const std::vector<uint64_t>& ins, Scope& scope, Iters& iters,
                           Queue& instances, Queue& witnesses, const uint32_t* weight) {
        Scope* ns = new_scope();
        try {
            for (size_t idx = 0; idx < ins.size(); idx++) set(*ns, idx + outs.size(), prog().copy(get(scope, ins[idx])));
            for (const auto& g : sub) ingest_gate(g, consts, *ns, iters, instances, witnesses, weight);
            for (size_t idx = 0; idx < outs.size(); idx++) set(scope, outs[idx], prog().copy(get(*ns, idx)));
        } catch (...) {
            release_scope(ns);
            throw;
        }
        release_scope(ns);
    }

    // ---- evaluator.rs:318-691 --------------------------------------------------------------------
    void ingest_gate(const ir::Gate& g, const std::vector<std::vector<uint8_t>>& consts, Scope& scope, Iters& iters,
                     Queue& instances, Queue& witnesses, const uint32_t* weight) {
        Program& p = prog();
        if (p.n_values() >= c->max_values) throw EvalErr{"zkb: resource limit exceeded (max_values)"};
        step();
        switch (g.type) {
            case ir::G_CONSTANT: {  // :345-348
                const auto& v = consts[g.const_idx];
                set(scope, g.w0, p.constant(v.data(), v.size()));
            } break;
            case ir::G_ASSERT_ZERO: {  // :350-364
                uint32_t w = get(scope, g.w0);
                uint32_t z = weight ? as_mul(*weight, w) : p.copy(w);
                p.assert_zero(z, g.w0);
                p.ir_gates++;
            } break;
            case ir::G_COPY:  // :366-370
                set(scope, g.w0, p.copy(get(scope, g.w1)));
                break;
            case ir::G_ADD: case ir::G_MUL: case ir::G_AND: case ir::G_XOR: {  // :372-384, 400-412
                uint32_t l = get(scope, g.w1);
                uint32_t r = get(scope, g.w2);
                uint32_t res = g.type == ir::G_ADD ? p.add(l, r) : g.type == ir::G_MUL ? p.multiply(l, r)
                               : g.type == ir::G_AND ? p.and_(l, r) : p.xor_(l, r);
                p.ir_gates++;
                set(scope, g.w0, res);
            } break;
            case ir::G_ADD_CONSTANT: case ir::G_MUL_CONSTANT: {  // :386-398
                uint32_t l = get(scope, g.w1);
                const auto& v = consts[g.const_idx];
                uint32_t res = g.type == ir::G_ADD_CONSTANT ? p.add_constant(l, v.data(), v.size()) : p.mul_constant(l, v.data(), v.size());
                p.ir_gates++;
                set(scope, g.w0, res);
            } break;
            case ir::G_NOT:  // :414-418
                p.ir_gates++;
                set(scope, g.w0, p.not_(get(scope, g.w1)));
                break;
            case ir::G_INSTANCE: {  // :420-427
                if (instances.empty()) throw EvalErr{"Not enough instance to consume"};
                uint32_t pos = instances.front();
                instances.pop_front();
                p.cb_count[CB_INSTANCE]++;
                if (pos + 1 > p.n_instance) p.n_instance = pos + 1;
                set(scope, g.w0, p.push_value(V_INSTANCE, 0, pos));
            } break;
            case ir::G_WITNESS: {  // :429-432 + PlaintextBackend::witness :944-946 (None => panic)
                if (witnesses.empty()) throw Fatal{"Missing witness value for PlaintextBackend"};
                uint32_t pos = witnesses.front();
                witnesses.pop_front();
                p.cb_count[CB_WITNESS]++;
                if (pos + 1 > p.n_witness) p.n_witness = pos + 1;
                set(scope, g.w0, p.push_value(V_WITNESS, 0, pos));
            } break;
            case ir::G_FREE: {  // :434-439
                uint64_t last = g.has_last ? g.w1 : g.w0;
                if (last > g.w0) step(last - g.w0);
                for (uint64_t w = g.w0; w <= last; w++) {
                    remove(scope, w);
                    if (w == UINT64_MAX) break;
                }
            } break;
            case ir::G_CALL: {  // :441-471
                auto it = known_functions.find(g.cx->name);
                if (it == known_functions.end()) throw EvalErr{"Unknown function"};
                const FunctionDecl& f = it->second;
                std::vector<uint64_t> eo, ei;
                expand_wirelist(g.cx->outputs, eo);
                expand_wirelist(g.cx->inputs, ei);
                check_arity(g.cx->name, f, eo.size(), ei.size());
                Iters fresh;  // named calls do NOT see the caller's iterators (:456)
                auto body = f.body;
                auto fc = f.consts;
                ingest_subcircuit(*body, *fc, eo, ei, scope, fresh, instances, witnesses, weight);
            } break;
            case ir::G_ANON_CALL: {  // :473-491
                std::vector<uint64_t> eo, ei;
                expand_wirelist(g.cx->outputs, eo);
                expand_wirelist(g.cx->inputs, ei);
                ingest_subcircuit(g.cx->body, consts, eo, ei, scope, iters, instances, witnesses, weight);
            } break;
            case ir::G_FOR: {  // :495-559
                const ir::Complex& cx = *g.cx;
                std::vector<uint64_t> eo, ei;
                for (uint64_t i = cx.first; i <= cx.last; i++) {
                    step();
                    iters_set(iters, cx.name, i);
                    if (!cx.body_is_anon) {
                        auto it = known_functions.find(cx.fn_name);
                        if (it == known_functions.end()) throw EvalErr{"Unknown function"};
                        const FunctionDecl& f = it->second;
                        eval_iterexpr_list(cx.it_outputs, iters, eo);
                        eval_iterexpr_list(cx.it_inputs, iters, ei);
                        check_arity(cx.fn_name, f, eo.size(), ei.size());
                        Iters fresh;
                        auto body = f.body;
                        auto fc = f.consts;
                        ingest_subcircuit(*body, *fc, eo, ei, scope, fresh, instances, witnesses, weight);
                    } else {
                        eval_iterexpr_list(cx.it_outputs, iters, eo);
                        eval_iterexpr_list(cx.it_inputs, iters, ei);
                        ingest_subcircuit(cx.body, consts, eo, ei, scope, iters, instances, witnesses, weight);
                    }
                    if (i == UINT64_MAX) break;
                }
                iters_remove(iters, cx.name);
            } break;
            case ir::G_SWITCH: {  // :563-688
                const ir::Complex& cx = *g.cx;
                uint64_t max_i = 0, max_w = 0;
                for (const auto& br : cx.branches) {  // :565-581
                    uint64_t ic, wc;
                    if (!br.is_anon) {
                        auto it = known_functions.find(br.name);
                        if (it == known_functions.end()) throw EvalErr{"Unknown function"};
                        ic = it->second.instance_nbr;
                        wc = it->second.witness_nbr;
                    } else {
                        ic = br.instance_count;
                        wc = br.witness_count;
                    }
                    max_i = std::max(max_i, ic);
                    max_w = std::max(max_w, wc);
                }
                // :586-591: the first `max` queued values are handed (cloned) to every branch
                Queue new_instances, new_witnesses;
                for (uint64_t k = 0, n = std::min<uint64_t>(instances.size(), max_i); k < n; k++) {
                    new_instances.push_back(instances.front());
                    instances.pop_front();
                }
                for (uint64_t k = 0, n = std::min<uint64_t>(witnesses.size(), max_w); k < n; k++) {
                    new_witnesses.push_back(witnesses.front());
                    witnesses.pop_front();
                }
                std::vector<uint64_t> eo, ei;
                expand_wirelist(cx.outputs, eo);  // :597
                std::vector<Scope*> branch_scopes;
                std::vector<uint32_t> weights;
                auto cleanup = [&]() {
                    for (auto s : branch_scopes) release_scope(s);
                    branch_scopes.clear();
                };
                try {
                    size_t nb = std::min(cx.cases.size(), cx.branches.size());  // zip
                    for (size_t k = 0; k < nb; k++) {  // :600-670
                        const auto& br = cx.branches[k];
                        uint32_t bw = compute_weight(consts[cx.cases[k]], get(scope, g.w0));
                        uint32_t wbw = weight ? as_mul(*weight, bw) : bw;
                        Scope* bs = new_scope();
                        branch_scopes.push_back(bs);
                        Queue qi = new_instances, qw = new_witnesses;
                        if (!br.is_anon) {
                            auto it = known_functions.find(br.name);
                            if (it == known_functions.end()) throw EvalErr{"Unknown function: " + br.name};
                            const FunctionDecl& f = it->second;
                            expand_wirelist(br.inputs, ei);
                            check_arity(br.name, f, eo.size(), ei.size());
                            for (uint64_t w : ei) bs->set(w, p.copy(get(scope, w)));  // HashMap::insert (:626-629)
                            Iters fresh;
                            auto body = f.body;
                            auto fc = f.consts;
                            ingest_subcircuit(*body, *fc, eo, ei, *bs, fresh, qi, qw, &wbw);
                        } else {
                            expand_wirelist(br.inputs, ei);
                            for (uint64_t w : ei) bs->set(w, p.copy(get(scope, w)));
                            ingest_subcircuit(br.subcircuit, consts, eo, ei, *bs, iters, qi, qw, &wbw);
                        }
                        weights.push_back(wbw);
                    }
                    for (uint64_t ow : eo) {  // :673-687
                        uint8_t zero = 0;
                        uint32_t acc = p.constant(&zero, 1);
                        for (size_t k = 0; k < branch_scopes.size(); k++) {
                            uint32_t ww = as_mul(get(*branch_scopes[k], ow), weights[k]);
                            acc = as_add(acc, ww);
                        }
                        set(scope, ow, acc);
                    }
                } catch (...) {
                    cleanup();
                    throw;
                }
                cleanup();
            } break;
            default:
                throw EvalErr{"No gate type"};
        }
    }

    // ---- evaluator.rs:232-303 ----------------------------------------------------------------------
    void ingest_values(const ir::Message& m) {
        modulus_le = m.header.field_characteristic;  // ingest_header: last header wins
        bool inst = m.type == ir::MSG_INSTANCE;
        auto& store = inst ? instance_values : witness_values;
        auto& q = inst ? instance_queue : witness_queue;
        for (const auto& v : m.values) {
            if (store.size() >= 0xFFFFFFF0u) throw EvalErr{"zkb: too many input values"};
            q.push_back((uint32_t)store.size());
            store.push_back(v);
        }
    }

    void ingest_relation(ir::Message& m) {
        modulus_le = m.header.field_characteristic;
        is_boolean = (m.gate_mask & ir::M_BOOL) == ir::M_BOOL;  // contains_feature(gate_mask, BOOL), :262
        std::string e;
        if (!prog().set_field(m.header.field_characteristic.data(), m.header.field_characteristic.size(), m.header.field_degree, e))
            throw EvalErr{e};
        c->is_boolean = is_boolean;
        if (!m.gates.empty()) verified_at_least_one_gate = true;
        auto consts = std::make_shared<std::vector<std::vector<uint8_t>>>(std::move(m.consts));
        for (auto& f : m.functions) {  // :273-284, later definitions overwrite
            FunctionDecl d;
            d.body = std::make_shared<std::vector<ir::Gate>>(std::move(f.body));
            d.consts = consts;
            d.instance_nbr = f.instance_count;
            d.witness_nbr = f.witness_count;
            d.output_count = f.output_count;
            d.input_count = f.input_count;
            known_functions[f.name] = std::move(d);
        }
        Iters iters;
        for (const auto& g : m.gates) ingest_gate(g, *consts, values, iters, instance_queue, witness_queue, nullptr);
    }

    // Evaluator::ingest_message, :213-230: errors latch, later messages are skipped
    int ingest_parsed(ir::Message& m) {
        if (fatal) return fail(ZKB_E_FATAL, err);
        if (evaluated) return fail(ZKB_E_ARG, "evaluator already finished (get_violations was called)");
        if (has_error) return ZKB_OK;
        try {
            if (m.type == ir::MSG_RELATION) ingest_relation(m);
            else ingest_values(m);
        } catch (const EvalErr& e) {
            has_error = true;
            found_error = e.msg;
            ctx_latch(c, e.msg);
        } catch (const Fatal& f) {
            fatal = true;
            return fail(ZKB_E_FATAL, f.msg);
        }
        return ZKB_OK;
    }

    int ingest_bytes(const uint8_t* buf, size_t len) {
        ir::Message m;
        std::string e;
        if (!ir::read_message(buf, len, m, e)) {
            // Evaluator::from_messages unwraps parse errors (:193): the reference aborts
            fatal = true;
            return fail(ZKB_E_FORMAT, e);
        }
        return ingest_parsed(m);
    }

    int fail(int code, const std::string& m) {
        err = m;
        return code;
    }

    // ---- get_violations, :199-208, with the deferred evaluation in the middle -------------------------
    int finish() {
        if (evaluated) return ZKB_OK;
        if (fatal) return fail(ZKB_E_FATAL, err);
        violations.clear();
        if (!verified_at_least_one_gate) violations.push_back("Did not receive any gate to verify.");
        std::string first_error;
        bool have_error = false;
        Program& p = prog();
        if (p.field_set && p.n_values() > 0) {
            if (!c->finalized) {
                c->live_values.clear();
                values.for_each([&](uint64_t, uint32_t v) { c->live_values.push_back(v); });
                int rc = ctx_finalize(c, 0);
                if (rc != ZKB_OK) return fail(rc, c->err);
            }
            // pack the queued streams: one statement = batch of 1
            size_t stride = (size_t)p.nlimb * 4;
            auto widen = [&](const std::vector<std::vector<uint8_t>>& vals) {
                for (const auto& v : vals) {
                    size_t n = v.size();
                    while (n > 0 && v[n - 1] == 0) n--;
                    stride = std::max(stride, (n + 3) / 4 * 4);
                }
            };
            widen(instance_values);
            widen(witness_values);
            auto pack = [&](const std::vector<std::vector<uint8_t>>& vals, uint32_t need) {
                std::vector<uint8_t> out((size_t)std::max<uint32_t>(need, 1) * stride, 0);
                for (uint32_t i = 0; i < need && i < vals.size(); i++) {
                    size_t n = vals[i].size();
                    while (n > 0 && vals[i][n - 1] == 0) n--;
                    memcpy(out.data() + (size_t)i * stride, vals[i].data(), n);
                }
                return out;
            };
            std::vector<uint8_t> ib = pack(instance_values, p.n_instance), wb = pack(witness_values, p.n_witness);
            zkb_verdict v;
            int rc = zkb_evaluate(c, ib.data(), 0, wb.data(), 0, (uint32_t)stride, 1, &v);
            if (rc != ZKB_OK) return fail(rc, c->err);
            if (v.first_fail_seq != UINT64_MAX) {
                uint64_t w = p.asserts[v.first_fail_seq].src_wire;
                first_error = "Wire_" + u64s(w) + " (may be weighted) should be 0, while it is not";  // :357-362
                have_error = true;
            }
        }
        if (!have_error && has_error) {
            first_error = found_error;
            have_error = true;
        }
        if (have_error) violations.push_back(first_error);
        evaluated = true;
        return ZKB_OK;
    }
};

// --------------------------------------------------------------------------------------------------
// Source: file discovery and ordering, rust/src/consumers/source.rs:64-89, 165-193
// --------------------------------------------------------------------------------------------------
static bool has_sieve_ext(const std::string& p) {
    size_t slash = p.find_last_of('/');
    std::string name = slash == std::string::npos ? p : p.substr(slash + 1);
    size_t dot = name.find_last_of('.');
    return dot != std::string::npos && dot > 0 && name.substr(dot + 1) == "sieve";
}

static int list_workspace_files(zkb_evaluator* ev, const char* const* paths, size_t n, std::vector<std::string>& out) {
    for (size_t i = 0; i < n; i++) {
        std::string p = paths[i];
        if (has_sieve_ext(p)) {
            out.push_back(p);
        } else if (p == "-") {
            return ev->fail(ZKB_E_UNSUPPORTED, "zkb: reading the statement from stdin is not supported");
        } else {
            DIR* d = opendir(p.c_str());
            if (!d) return ev->fail(ZKB_E_ARG, "cannot read directory " + p);
            while (dirent* e = readdir(d)) {
                std::string name = e->d_name;
                std::string full = p + (p.size() && p.back() == '/' ? "" : "/") + name;
                if (has_sieve_ext(full)) out.push_back(full);
            }
            closedir(d);
        }
    }
    // from_filenames: lexical sort, then a STABLE sort on instance < witness < relation < other
    std::sort(out.begin(), out.end());
    auto key = [](const std::string& p) {
        size_t slash = p.find_last_of('/');
        std::string name = slash == std::string::npos ? p : p.substr(slash + 1);
        if (name.find("instance") != std::string::npos) return 0;
        if (name.find("witness") != std::string::npos) return 1;
        if (name.find("relation") != std::string::npos) return 3;
        return 4;
    };
    std::stable_sort(out.begin(), out.end(), [&](const std::string& a, const std::string& b) { return key(a) < key(b); });
    return ZKB_OK;
}

extern "C" zkb_evaluator* zkb_evaluator_create(zkb_ctx* backend) { return new zkb_evaluator(backend); }
extern "C" void zkb_evaluator_destroy(zkb_evaluator* ev) { delete ev; }
extern "C" const char* zkb_evaluator_last_error(zkb_evaluator* ev) { return ev->err.c_str(); }

extern "C" int zkb_evaluator_ingest_message(zkb_evaluator* ev, const uint8_t* buf, size_t len) { return ev->ingest_bytes(buf, len); }

extern "C" int zkb_evaluator_ingest_buffer(zkb_evaluator* ev, const uint8_t* buf, size_t len) {
    std::vector<std::pair<size_t, size_t>> msgs;
    ir::split_messages(buf, len, msgs);
    for (auto& m : msgs) {
        int rc = ev->ingest_bytes(buf + m.first, m.second);
        if (rc != ZKB_OK) return rc;
    }
    return ZKB_OK;
}

extern "C" int zkb_evaluator_ingest_paths(zkb_evaluator* ev, const char* const* paths, size_t n_paths) {
    std::vector<std::string> files;
    int rc = list_workspace_files(ev, paths, n_paths, files);
    if (rc != ZKB_OK) return rc;
    for (const auto& f : files) {
        FILE* fp = fopen(f.c_str(), "rb");
        if (!fp) {
            fprintf(stderr, "Warning: failed to open file %s\n", f.c_str());  // source.rs:132
            continue;
        }
        std::vector<uint8_t> data;
        uint8_t chunk[1 << 16];
        size_t got;
        while ((got = fread(chunk, 1, sizeof chunk, fp)) > 0) data.insert(data.end(), chunk, chunk + got);
        fclose(fp);
        rc = zkb_evaluator_ingest_buffer(ev, data.data(), data.size());
        if (rc != ZKB_OK) return rc;
    }
    return ZKB_OK;
}

extern "C" int zkb_evaluator_get_violations(zkb_evaluator* ev, size_t* n) {
    int rc = ev->finish();
    if (rc != ZKB_OK) return rc;
    *n = ev->violations.size();
    return ZKB_OK;
}

extern "C" const char* zkb_evaluator_violation(zkb_evaluator* ev, size_t i) {
    return i < ev->violations.size() ? ev->violations[i].c_str() : nullptr;
}

extern "C" int zkb_evaluator_get_wire(zkb_evaluator* ev, uint64_t wire_id, uint8_t* out, size_t cap, size_t* len) {
    int rc = ev->finish();
    if (rc != ZKB_OK) return rc;
    uint32_t v = ev->values.get(wire_id);
    if (v == Scope::kNone) return ev->fail(ZKB_E_SEMANTIC, "No value given for wire_" + u64s(wire_id));  // :750-752, 787-791
    zkb_wire w = v;
    memset(out, 0, cap);
    rc = zkb_read_values(ev->c, 0, &w, 1, out, cap);
    if (rc != ZKB_OK) return ev->fail(rc, ev->c->err);
    size_t n = cap;
    while (n > 1 && out[n - 1] == 0) n--;
    if (len) *len = n;
    return ZKB_OK;
}

extern "C" int zkb_evaluator_lookup(zkb_evaluator* ev, uint64_t wire_id, zkb_wire* out) {
    uint32_t v = ev->values.get(wire_id);
    if (v == Scope::kNone) return ev->fail(ZKB_E_SEMANTIC, "No value given for wire_" + u64s(wire_id));
    *out = v;
    return ZKB_OK;
}
