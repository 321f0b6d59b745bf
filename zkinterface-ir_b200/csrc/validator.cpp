// zkb_validator — mirror of `Validator` (rust/src/consumers/validator.rs:68-829): the semantic / syntactic
// checks `zki_sieve validate` and `valid-eval-metrics` run (cli.rs:302-313, 333-363).  Host only.
//
// Control flow, state and violation texts follow the reference check for check (the texts are the drop-in
// surface: callers print and compare them).  Differences, all outside the reference's own behaviour:
//   - `probably_prime(n, 10)` of crate num-bigint-dig is a probabilistic test; here Miller-Rabin over the first
//     24 primes (same answer except with negligible probability);
//   - the two regular expressions (validator.rs:23-25) are hand-written matchers over ASCII classes
//     (`\d` = [0-9], `\w` = [A-Za-z0-9_]; the regex crate's classes also take non-ASCII digits / letters);
//   - loops are bounded by zkb_validator_set_limits instead of running until memory is exhausted.
#include <dirent.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <unordered_map>
#include <unordered_set>

#include "../../include/zkb.h"
#include "bigu.h"
#include "file_bytes.h"
#include "ir.h"

using namespace zkb;

namespace {

const char* const kNamesRegex = "^[a-zA-Z_][\\w]*(?:(?:\\.|:{2})[a-zA-Z_][\\w]*)*$";

struct Fatal {
    std::string msg;
};

std::string u64s(uint64_t v) { return std::to_string((unsigned long long)v); }

std::string big_dec(const BigU& v) { return v.to_dec(); }  // Display of a BigUint

std::string debug_bytes(const std::vector<uint8_t>& v) {  // `{:?}` of a Vec<u8>
    std::string s = "[";
    for (size_t i = 0; i < v.size(); i++) {
        if (i) s += ", ";
        s += std::to_string((unsigned)v[i]);
    }
    return s + "]";
}

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    auto ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
    while (a < b && ws((unsigned char)s[a])) a++;
    while (b > a && ws((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

bool is_digit(char c) { return c >= '0' && c <= '9'; }
bool is_word(char c) { return is_digit(c) || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
bool is_name_start(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }

// ^\d+.\d+.\d+$ with `.` = any character but a line feed.  Backtracking over the three greedy runs.
bool match_version_from(const std::string& s, size_t pos, int part) {
    size_t end = pos;
    while (end < s.size() && is_digit(s[end])) end++;
    if (end == pos) return false;
    if (part == 2) return end == s.size();
    // `\d+` may give characters back: the separator `.` can itself be a digit
    for (size_t e = end; e > pos; e--) {
        if (e >= s.size() || s[e] == '\n') continue;
        size_t next = e + 1;  // `.` is one CHARACTER: skip the continuation bytes of a multi-byte one
        while (next < s.size() && ((unsigned char)s[next] & 0xC0) == 0x80) next++;
        if (match_version_from(s, next, part + 1)) return true;
    }
    return false;
}
bool match_version(const std::string& s) { return match_version_from(s, 0, 0); }

// ^[a-zA-Z_][\w]*(?:(?:\.|:{2})[a-zA-Z_][\w]*)*$
bool match_name(const std::string& s) {
    size_t i = 0;
    auto ident = [&]() {
        if (i >= s.size() || !is_name_start(s[i])) return false;
        i++;
        while (i < s.size() && is_word(s[i])) i++;
        return true;
    };
    if (!ident()) return false;
    while (i < s.size()) {
        if (s[i] == '.') i += 1;
        else if (s[i] == ':' && i + 1 < s.size() && s[i + 1] == ':') i += 2;
        else return false;
        if (!ident()) return false;
    }
    return true;
}

BigU powmod(BigU base, const BigU& exp, const BigU& m) {
    BigU r(1);
    base = base.mod(m);
    for (size_t i = exp.bits(); i-- > 0;) {
        r = BigU::mulmod(r, r, m);
        if (exp.bit(i)) r = BigU::mulmod(r, base, m);
    }
    return r;
}

bool is_probably_prime(const std::vector<uint8_t>& value) {  // structs/value.rs:53-56
    static const uint32_t kPrimes[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89};
    BigU n = BigU::from_bytes_le(value.data(), value.size());
    if (n.is_zero() || n.is_one()) return false;
    for (uint32_t p : kPrimes) {
        if (n.mod(BigU(p)).is_zero()) return n == BigU(p);
    }
    BigU nm1 = n;
    nm1.sub(BigU(1));
    BigU d = nm1;
    size_t s = 0;
    while (!d.bit(0)) {
        d.shr1();
        s++;
    }
    for (uint32_t a : kPrimes) {
        BigU x = powmod(BigU(a), d, n);
        if (x.is_one() || x == nm1) continue;
        bool composite = true;
        for (size_t r = 1; r < s; r++) {
            x = BigU::mulmod(x, x, n);
            if (x == nm1) {
                composite = false;
                break;
            }
        }
        if (composite) return false;
    }
    return true;
}

// BTreeSet<WireId>: dense bitmap for small ids, hash set beyond
class LiveSet {  // HashSet<WireId> (validator.rs:74): a bitmap while ids stay within 64 x the insertions made, a hash set beyond,
                 // so memory follows the work done and not the largest id a gate names
public:
    static constexpr uint64_t kDense = 1ull << 36;
    bool contains(uint64_t id) const {
        if ((id >> 6) < bits_.size() && ((bits_[id >> 6] >> (id & 63)) & 1)) return true;
        return !sparse_.empty() && sparse_.count(id) != 0;
    }
    void insert(uint64_t id) {
        if (contains(id)) return;
        inserts_++;
        if ((id >> 6) >= bits_.size() && id < kDense && id <= 64 * (inserts_ + 1024))
            bits_.resize(std::max<size_t>((id >> 6) + 1, bits_.size() * 2), 0);
        if ((id >> 6) < bits_.size()) bits_[id >> 6] |= 1ull << (id & 63);
        else sparse_.insert(id);
    }
    bool erase(uint64_t id) {
        if ((id >> 6) < bits_.size() && ((bits_[id >> 6] >> (id & 63)) & 1)) {
            bits_[id >> 6] &= ~(1ull << (id & 63));
            return true;
        }
        return !sparse_.empty() && sparse_.erase(id) != 0;
    }
    size_t count() const {
        size_t n = sparse_.size();
        for (uint64_t w : bits_) n += (size_t)__builtin_popcountll(w);
        return n;
    }

private:
    std::vector<uint64_t> bits_;
    std::unordered_set<uint64_t> sparse_;
    uint64_t inserts_ = 0;
};

struct FnSig {
    uint64_t output_count, input_count, instance_count, witness_count;
};
using Iters = std::vector<std::pair<std::string, uint64_t>>;

struct Shared {  // what the reference shares between a validator and its sub-validators through Rc<RefCell<..>>
    std::unordered_map<std::string, FnSig> known_functions;
    uint64_t steps = 0, max_steps = 1ull << 40;
};

}  // namespace

struct zkb_validator {
    // validator.rs:68-87
    bool as_prover = false;
    uint64_t instance_queue_len = 0, witness_queue_len = 0;
    LiveSet live_wires;
    bool got_header = false;
    uint16_t gate_set = 0, features = 0;
    std::string header_version;
    BigU field_characteristic;
    uint64_t field_degree = 0;
    std::shared_ptr<Shared> shared = std::make_shared<Shared>();
    std::shared_ptr<Iters> known_iterators = std::make_shared<Iters>();
    std::vector<std::string> violations;

    std::string err;
    bool fatal = false, finished = false;
    const std::vector<std::vector<uint8_t>>* consts = nullptr;  // constant table of the message being ingested

    void violate(const std::string& m) { violations.push_back(m); }
    void step(uint64_t n = 1) {
        shared->steps += n;
        if (shared->steps > shared->max_steps || shared->steps < n) throw Fatal{"zkb: resource limit exceeded (max_steps)"};
    }

    // ---- helpers, validator.rs:740-829 --------------------------------------------------------------
    bool is_defined(uint64_t id) const { return live_wires.contains(id); }
    void declare(uint64_t id) { live_wires.insert(id); }
    void remove(uint64_t id) {
        if (!live_wires.erase(id))
            violate("The variable " + u64s(id) + " is being freed, but was not defined previously, or has been already freed");
    }
    void consume_instance(uint64_t how_many) {
        if (instance_queue_len >= how_many) {
            instance_queue_len -= how_many;
        } else {
            instance_queue_len = 0;
            violate("Not enough Instance value to consume.");
        }
    }
    void consume_witness(uint64_t how_many) {
        if (!as_prover) return;
        if (witness_queue_len >= how_many) {
            witness_queue_len -= how_many;
        } else {
            witness_queue_len = 0;
            violate("Not enough Witness value to consume.");
        }
    }
    void ensure_defined_and_set(uint64_t id) {
        if (!is_defined(id)) {
            if (as_prover) violate("The wire " + u64s(id) + " is used but was not assigned a value, or has been freed already.");
            declare(id);
        }
    }
    void ensure_undefined_and_set(uint64_t id) {
        if (is_defined(id)) violate("The wire " + u64s(id) + " has already been initialized before. This violates the SSA property.");
        declare(id);
    }
    template <class F>
    void ensure_value_in_field(const std::vector<uint8_t>& value, F name) {
        if (value.empty()) violate("The " + name() + " is empty.");
        BigU v = BigU::from_bytes_le(value.data(), value.size());
        if (v >= field_characteristic)
            violate("The " + name() + " cannot be represented in the field specified in Header (" + big_dec(v) + " >= " +
                    big_dec(field_characteristic) + ").");
    }
    void ensure_allowed_gate(const char* name, uint16_t mask) {
        if ((gate_set & mask) != mask) violate(std::string("The gate ") + name + " is not allowed in this circuit.");
    }
    void ensure_allowed_feature(const char* name, uint16_t mask) {
        if ((features & mask) != mask) violate(std::string("The feature ") + name + " is not allowed in this circuit.");
    }

    // ---- structs/wire.rs:179-203 (errors become violations, the list becomes empty) ---------------------
    std::vector<uint64_t> expand(const ir::WireList& wl) {
        std::vector<uint64_t> out;
        for (const auto& e : wl) {
            if (!e.is_range) {
                out.push_back(e.first);
            } else {
                if (e.last <= e.first) {
                    violate("In WireRange, last WireId (" + u64s(e.last) + ") must be strictly greater than first WireId (" +
                            u64s(e.first) + ").");
                    return {};
                }
                step(e.last - e.first);
                for (uint64_t w = e.first; w <= e.last; w++) out.push_back(w);
            }
        }
        return out;
    }
    // ---- structs/iterators.rs:349-403 (errors are panics) ----------------------------------------------
    uint64_t eval_iterexpr(const ir::IterExpr& e) {
        switch (e.type) {
            case 1: return e.value;
            case 2:
                for (size_t i = known_iterators->size(); i-- > 0;)
                    if ((*known_iterators)[i].first == e.name) return (*known_iterators)[i].second;
                throw Fatal{"Unknown iterator name " + e.name};
            case 3: return eval_iterexpr(*e.l) + eval_iterexpr(*e.r);
            case 4: return eval_iterexpr(*e.l) - eval_iterexpr(*e.r);
            case 5: return eval_iterexpr(*e.l) * eval_iterexpr(*e.r);
            case 6:
                if (e.value == 0) throw Fatal{"attempt to divide by zero"};
                return eval_iterexpr(*e.l) / e.value;
        }
        throw Fatal{"Unknown Iterator Expression type"};
    }
    std::vector<uint64_t> iterexprs(const ir::IterExprList& l) {
        std::vector<uint64_t> out;
        for (const auto& el : l) {
            uint64_t a = eval_iterexpr(el.first);
            if (!el.is_range) {
                out.push_back(a);
            } else {
                uint64_t b = eval_iterexpr(el.last);
                if (b >= a) step(b - a);
                for (uint64_t w = a; w <= b && b >= a; w++) {
                    out.push_back(w);
                    if (w == UINT64_MAX) break;
                }
            }
        }
        return out;
    }

    // ---- validator.rs:162-202 ---------------------------------------------------------------------------
    void ingest_header(const ir::Header& h) {
        BigU fc = BigU::from_bytes_le(h.field_characteristic.data(), h.field_characteristic.size());
        if (got_header) {
            if (!(field_characteristic == fc)) violate("The field_characteristic field is not consistent across headers.");
            if (field_degree != h.field_degree) violate("The field_degree is not consistent across headers.");
            if (header_version != h.version) violate("The profile version is not consistent across headers.");
        } else {
            got_header = true;
            field_characteristic = fc;
            if (fc.is_zero() || fc.is_one()) violate("The field_characteristic should be > 1");
            if (!is_probably_prime(h.field_characteristic)) violate("The field_characteristic should be a prime.");
            field_degree = h.field_degree;
            if (field_degree != 1) violate("field_degree must be = 1");
            if (!match_version(trim(h.version)))
                violate("The profile version should match the following format <major>.<minor>.<patch>.");
            header_version = h.version;
        }
    }

    // ---- validator.rs:204-289 ---------------------------------------------------------------------------
    void ingest_message(const ir::Message& m) {
        consts = &m.consts;
        if (m.type == ir::MSG_INSTANCE) {
            ingest_header(m.header);
            for (const auto& v : m.values) ensure_value_in_field(v, [&]() { return "instance value " + debug_bytes(v); });
            instance_queue_len += m.values.size();
        } else if (m.type == ir::MSG_WITNESS) {
            if (!as_prover) violate("As verifier, got an unexpected Witness message.");
            ingest_header(m.header);
            for (const auto& v : m.values) ensure_value_in_field(v, [&]() { return "witness value " + debug_bytes(v); });
            witness_queue_len += m.values.size();
        } else {
            ingest_header(m.header);
            gate_set = m.gate_mask;
            if ((gate_set & ir::M_BOOL) == ir::M_BOOL && (gate_set & ir::M_ARITH) == ir::M_ARITH)
                violate("Cannot mix arithmetic and boolean gates");
            if ((gate_set & ir::M_BOOL) == ir::M_BOOL) {
                if (!(field_characteristic == BigU(2))) violate("With boolean profile the field characteristic can only be 2.");
            }
            features = m.feat_mask;
            for (const auto& f : m.functions) {
                ensure_allowed_feature("@function", ir::M_FUNCTION);
                if (!match_name(trim(f.name)))
                    violate("The function name (" + f.name + ") should match the proper format (" + kNamesRegex + ").");
                if (shared->known_functions.count(f.name)) {
                    violate("A function with the name '" + f.name + "' already exists");
                    continue;
                }
                shared->known_functions[f.name] = FnSig{f.output_count, f.input_count, f.instance_count, f.witness_count};
                ingest_subcircuit(f.body, f.output_count, f.input_count, f.instance_count, f.witness_count, false);
            }
            for (const auto& g : m.gates) ingest_gate(g);
        }
    }

    // validator.rs:649-673; false stands for Err (unknown function)
    bool ingest_call(const std::string& name, size_t n_out, size_t n_in, uint64_t& ic, uint64_t& wc) {
        ic = wc = 0;
        auto it = shared->known_functions.find(name);
        if (it == shared->known_functions.end()) {
            violate("Unknown Function gate " + name);
            return false;
        }
        if (it->second.output_count != n_out) violate("Call: number of output wires mismatch.");
        if (it->second.input_count != n_in) violate("Call: number of input wires mismatch.");
        ic = it->second.instance_count;
        wc = it->second.witness_count;
        return true;
    }

    // validator.rs:684-738
    void ingest_subcircuit(const std::vector<ir::Gate>& sub, uint64_t output_count, uint64_t input_count, uint64_t instance_count,
                           uint64_t witness_count, bool use_same_scope) {
        zkb_validator cur;
        cur.as_prover = as_prover;
        cur.instance_queue_len = instance_count;
        cur.witness_queue_len = as_prover ? witness_count : 0;
        cur.got_header = got_header;
        cur.gate_set = gate_set;
        cur.features = features;
        cur.header_version = header_version;
        cur.field_characteristic = field_characteristic;
        cur.field_degree = field_degree;
        cur.shared = shared;
        cur.known_iterators = use_same_scope ? known_iterators : std::make_shared<Iters>();
        cur.consts = consts;
        step(input_count);
        for (uint64_t w = output_count; w < output_count + input_count; w++) cur.live_wires.insert(w);
        for (const auto& g : sub) cur.ingest_gate(g);
        step(output_count);
        for (uint64_t w = 0; w < output_count; w++) cur.ensure_defined_and_set(w);
        for (auto& v : cur.violations) violations.push_back(std::move(v));
        if (cur.instance_queue_len != 0) violate("The subcircuit has not consumed all the instance variables it should have.");
        if (cur.witness_queue_len != 0) violate("The subcircuit has not consumed all the witness variables it should have.");
    }

    // validator.rs:291-642
    void ingest_gate(const ir::Gate& g) {
        step();
        switch (g.type) {
            case ir::G_CONSTANT:
                ensure_value_in_field((*consts)[g.const_idx], []() { return std::string("Gate::Constant constant"); });
                ensure_undefined_and_set(g.w0);
                break;
            case ir::G_ASSERT_ZERO: ensure_defined_and_set(g.w0); break;
            case ir::G_COPY:
                ensure_defined_and_set(g.w1);
                ensure_undefined_and_set(g.w0);
                break;
            case ir::G_ADD: case ir::G_MUL: case ir::G_AND: case ir::G_XOR:
                if (g.type == ir::G_ADD) ensure_allowed_gate("@add", ir::M_ADD);
                else if (g.type == ir::G_MUL) ensure_allowed_gate("@mul", ir::M_MUL);
                else if (g.type == ir::G_AND) ensure_allowed_gate("@and", ir::M_AND);
                else ensure_allowed_gate("@xor", ir::M_XOR);
                ensure_defined_and_set(g.w1);
                ensure_defined_and_set(g.w2);
                ensure_undefined_and_set(g.w0);
                break;
            case ir::G_ADD_CONSTANT: case ir::G_MUL_CONSTANT: {
                const bool add = g.type == ir::G_ADD_CONSTANT;
                ensure_allowed_gate(add ? "@addc" : "@mulc", add ? ir::M_ADDC : ir::M_MULC);
                ensure_value_in_field((*consts)[g.const_idx],
                                      [&]() { return std::string(add ? "Gate::AddConstant_" : "Gate::MulConstant_") + u64s(g.w0); });
                ensure_defined_and_set(g.w1);
                ensure_undefined_and_set(g.w0);
            } break;
            case ir::G_NOT:
                ensure_allowed_gate("@not", ir::M_NOT);
                ensure_defined_and_set(g.w1);
                ensure_undefined_and_set(g.w0);
                break;
            case ir::G_INSTANCE:
                declare(g.w0);
                consume_instance(1);
                break;
            case ir::G_WITNESS:
                declare(g.w0);
                consume_witness(1);
                break;
            case ir::G_FREE: {
                if (g.has_last && g.w1 <= g.w0)
                    violate("For Free gates, last WireId (" + u64s(g.w1) + ") must be strictly greater than first WireId (" +
                            u64s(g.w0) + ").");
                uint64_t last = g.has_last ? g.w1 : g.w0;
                if (last > g.w0) step(last - g.w0);
                for (uint64_t w = g.w0; w <= last; w++) {
                    ensure_defined_and_set(w);
                    remove(w);
                    if (w == UINT64_MAX) break;
                }
            } break;
            case ir::G_ANON_CALL: {
                ensure_allowed_feature("@anoncall", ir::M_FUNCTION);
                std::vector<uint64_t> eo = expand(g.cx->outputs), ei = expand(g.cx->inputs);
                for (uint64_t w : ei) ensure_defined_and_set(w);
                ingest_subcircuit(g.cx->body, eo.size(), ei.size(), g.cx->instance_count, g.cx->witness_count, true);
                consume_instance(g.cx->instance_count);
                consume_witness(g.cx->witness_count);
                for (uint64_t w : eo) ensure_undefined_and_set(w);
            } break;
            case ir::G_CALL: {
                ensure_allowed_feature("@call", ir::M_FUNCTION);
                std::vector<uint64_t> eo = expand(g.cx->outputs), ei = expand(g.cx->inputs);
                for (uint64_t w : ei) ensure_defined_and_set(w);
                uint64_t ic, wc;
                ingest_call(g.cx->name, eo.size(), ei.size(), ic, wc);
                consume_instance(ic);
                consume_witness(wc);
                for (uint64_t w : eo) ensure_undefined_and_set(w);
            } break;
            case ir::G_SWITCH: {
                const ir::Complex& cx = *g.cx;
                ensure_allowed_feature("@switch", ir::M_SWITCH);
                ensure_defined_and_set(g.w0);
                if (cx.cases.size() != cx.branches.size())
                    violate("Gate::Switch: The number of cases value does not match the number of branches.");
                if (cx.cases.empty()) {
                    if (!cx.outputs.empty()) violate("Switch: no case given while non-empty list of output wires.");
                    return;
                }
                std::vector<BigU> seen;
                for (uint32_t ci : cx.cases) {
                    const auto& cv = (*consts)[ci];
                    BigU v = BigU::from_bytes_le(cv.data(), cv.size());
                    ensure_value_in_field(cv, [&]() { return "Gate::Switch case value: " + big_dec(v); });
                    bool dup = false;
                    for (const auto& s : seen) dup = dup || s == v;
                    if (!dup) seen.push_back(v);
                }
                if (seen.size() != cx.cases.size()) violate("Gate::Switch: The cases values contain duplicates.");
                uint64_t max_i = 0, max_w = 0;
                std::vector<uint64_t> eo = expand(cx.outputs);
                for (const auto& br : cx.branches) {
                    uint64_t ic = 0, wc = 0;
                    std::vector<uint64_t> ei = expand(br.inputs);
                    for (uint64_t w : ei) ensure_defined_and_set(w);
                    if (!br.is_anon) {
                        ingest_call(br.name, eo.size(), ei.size(), ic, wc);
                    } else {
                        ingest_subcircuit(br.subcircuit, eo.size(), ei.size(), br.instance_count, br.witness_count, true);
                        ic = br.instance_count;
                        wc = br.witness_count;
                    }
                    max_i = std::max(max_i, ic);
                    max_w = std::max(max_w, wc);
                }
                consume_instance(max_i);
                consume_witness(max_w);
                for (uint64_t w : eo) ensure_undefined_and_set(w);
            } break;
            case ir::G_FOR: {
                const ir::Complex& cx = *g.cx;
                ensure_allowed_feature("@for", ir::M_FOR);
                if (cx.last < cx.first) {
                    violate("In a For loop, the end value (" + u64s(cx.last) + ") must be strictly greater than the start value (" +
                            u64s(cx.first) + ").");
                    return;
                }
                for (const auto& kv : *known_iterators)
                    if (kv.first == cx.name) {
                        violate("Iterator already used in this context.");
                        return;
                    }
                if (!match_name(cx.name))
                    violate("The iterator name (" + cx.name + ") should match the following format (" + kNamesRegex + ").");
                for (uint64_t i = cx.first; i <= cx.last; i++) {
                    step();
                    // HashMap::insert on the shared map: a nested loop may have removed / re-added the name
                    bool found = false;
                    for (auto& kv : *known_iterators)
                        if (kv.first == cx.name) {
                            kv.second = i;
                            found = true;
                        }
                    if (!found) known_iterators->push_back({cx.name, i});
                    std::vector<uint64_t> eo = iterexprs(cx.it_outputs), ei = iterexprs(cx.it_inputs);
                    for (uint64_t w : ei) ensure_defined_and_set(w);
                    uint64_t ic = 0, wc = 0;
                    if (!cx.body_is_anon) {
                        ingest_call(cx.fn_name, eo.size(), ei.size(), ic, wc);
                    } else {
                        ingest_subcircuit(cx.body, eo.size(), ei.size(), cx.instance_count, cx.witness_count, true);
                        ic = cx.instance_count;
                        wc = cx.witness_count;
                    }
                    for (uint64_t w : eo) ensure_undefined_and_set(w);
                    consume_instance(ic);
                    consume_witness(wc);
                    if (i == UINT64_MAX) break;
                }
                for (size_t k = 0; k < known_iterators->size(); k++)
                    if ((*known_iterators)[k].first == cx.name) {
                        known_iterators->erase(known_iterators->begin() + k);
                        break;
                    }
                for (uint64_t w : expand(cx.outputs)) ensure_defined_and_set(w);
            } break;
            default: throw Fatal{"No gate type"};
        }
    }

    int fail(int code, const std::string& m) {
        err = m;
        return code;
    }

    int ingest_bytes(const uint8_t* buf, size_t len) {
        if (fatal) return fail(ZKB_E_FATAL, err);
        if (finished) return fail(ZKB_E_ARG, "validator already finished (get_violations was called)");
        ir::Message m;
        std::string e;
        if (!ir::read_message(buf, len, m, e)) {  // main_validate propagates `msg?` (cli.rs:304-305)
            fatal = true;
            return fail(ZKB_E_FORMAT, e);
        }
        try {
            ingest_message(m);
        } catch (const Fatal& f) {
            fatal = true;
            return fail(ZKB_E_FATAL, f.msg);
        }
        return ZKB_OK;
    }

    // get_violations, validator.rs:136-144
    int finish() {
        if (fatal) return fail(ZKB_E_FATAL, err);
        if (finished) return ZKB_OK;
        if (instance_queue_len > 0) violate("Too many Instance values (" + u64s(instance_queue_len) + " not consumed)");
        if (as_prover && witness_queue_len > 0) violate("Too many Witness values (" + u64s(witness_queue_len) + " not consumed)");
        finished = true;
        return ZKB_OK;
    }
};

extern "C" zkb_validator* zkb_validator_create(int as_prover) {
    zkb_validator* v = new zkb_validator();
    v->as_prover = as_prover != 0;
    return v;
}
extern "C" void zkb_validator_destroy(zkb_validator* v) { delete v; }
extern "C" const char* zkb_validator_last_error(zkb_validator* v) { return v->err.c_str(); }
extern "C" int zkb_validator_set_limits(zkb_validator* v, uint64_t max_steps) {
    if (max_steps) v->shared->max_steps = max_steps;
    return ZKB_OK;
}
extern "C" int zkb_validator_ingest_message(zkb_validator* v, const uint8_t* buf, size_t len) { return v->ingest_bytes(buf, len); }
extern "C" int zkb_validator_ingest_buffer(zkb_validator* v, const uint8_t* buf, size_t len) {
    std::vector<std::pair<size_t, size_t>> msgs;
    ir::split_messages(buf, len, msgs);
    for (auto& m : msgs) {
        int rc = v->ingest_bytes(buf + m.first, m.second);
        if (rc != ZKB_OK) return rc;
    }
    return ZKB_OK;
}

namespace zkb {
// shared with evaluator.cpp: Source::from_dirs_and_files ordering (source.rs:64-89, 165-193)
int list_workspace_files(const char* const* paths, size_t n, std::vector<std::string>& out, std::string& err);
}  // namespace zkb

extern "C" int zkb_validator_ingest_paths(zkb_validator* v, const char* const* paths, size_t n_paths) {
    std::vector<std::string> files;
    std::string e;
    int rc = zkb::list_workspace_files(paths, n_paths, files, e);
    if (rc != ZKB_OK) return v->fail(rc, e);
    for (const auto& f : files) {
        zkb::FileBytes data;
        if (!data.open(f)) {
            fprintf(stderr, "Warning: failed to open file %s\n", f.c_str());  // source.rs:132
            continue;
        }
        rc = zkb_validator_ingest_buffer(v, data.data, data.size);
        if (rc != ZKB_OK) return rc;
    }
    return ZKB_OK;
}
extern "C" int zkb_validator_get_violations(zkb_validator* v, size_t* n) {
    int rc = v->finish();
    if (rc != ZKB_OK) return rc;
    *n = v->violations.size();
    return ZKB_OK;
}
extern "C" const char* zkb_validator_violation(zkb_validator* v, size_t i) {
    return i < v->violations.size() ? v->violations[i].c_str() : nullptr;
}
// get_strict_violations / how_many_violations (validator.rs:146-152): without the end-of-statement checks
extern "C" size_t zkb_validator_how_many_violations(zkb_validator* v) { return v->violations.size(); }
// wires still live at the end (the reference prints "WARNING: few variables were not freed.", validator.rs:139-141)
extern "C" uint64_t zkb_validator_live_wires(zkb_validator* v) { return v->live_wires.count(); }
