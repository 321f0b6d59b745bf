// FlatBuffers writer for sieve_ir.fbs — the `build` half of rust/src/structs/{gates.rs:263-679,
// wire.rs:23-143, iterators.rs:112-312, function.rs:48-119,174-267, header.rs:58-76, value.rs:31-51,
// relation.rs:86-117, instance.rs:50-76, witness.rs:50-76} without flatc or the flatbuffers runtime.
//
// A message is built back to front like every FlatBuffers builder does (children before parents, the
// buffer grows towards lower addresses), with vtable de-duplication: a flat relation costs 76 bytes per
// two-input gate (three 16-byte Wire tables, gate table, Directive table, vector slot).
// Output is size-prefixed with the "siev" file identifier (sieve_ir.fbs, bottom).
#include <string.h>

#include <map>

#include "ir.h"

namespace zkb {
namespace ir {

namespace {

class Builder {
public:
    using Off = uint32_t;  // distance of an object's first byte from the END of the buffer

    Builder() : buf_(1024), head_(1024) {}

    Off size() const { return (Off)(buf_.size() - head_); }

    // ---- scalars / raw bytes -------------------------------------------------------------------
    void pad(size_t n) {
        grow(n);
        head_ -= n;
        memset(&buf_[head_], 0, n);
    }
    // make the write position a multiple of `align` AFTER `additional` more bytes are written
    void prep(size_t align, size_t additional) {
        if (align > minalign_) minalign_ = align;
        size_t after = (size_t)size() + additional;
        pad((align - after % align) % align);
    }
    template <class T>
    void push(T v) {
        grow(sizeof(T));
        head_ -= sizeof(T);
        memcpy(&buf_[head_], &v, sizeof(T));
    }
    void push_bytes(const void* p, size_t n) {
        grow(n);
        head_ -= n;
        if (n) memcpy(&buf_[head_], p, n);
    }
    // a uoffset field referring to `target`: value = (address of target) - (address of the field)
    void push_ref(Off target) {
        prep(4, 0);
        push<uint32_t>(size() + 4 - target);
    }

    // ---- vectors / strings ----------------------------------------------------------------------
    Off bytes_vector(const uint8_t* p, size_t n) {
        prep(4, n);
        push_bytes(p, n);
        push<uint32_t>((uint32_t)n);
        return size();
    }
    Off string(const std::string& s) {
        prep(4, s.size() + 1);
        push<uint8_t>(0);
        push_bytes(s.data(), s.size());
        push<uint32_t>((uint32_t)s.size());
        return size();
    }
    Off offset_vector(const std::vector<Off>& elems) {
        prep(4, elems.size() * 4);
        for (size_t i = elems.size(); i-- > 0;) push_ref(elems[i]);
        push<uint32_t>((uint32_t)elems.size());
        return size();
    }

    // ---- tables ----------------------------------------------------------------------------------
    // Fields are declared largest first (u64, then offsets, then bytes) so no padding is needed between them.
    void start_table() {
        n_fields_ = 0;
        table_end_ = size();
    }
    void field_u64(int slot, uint64_t v) {
        if (v == 0) return;  // default values are omitted, as flatc-generated builders do
        prep(8, 0);
        push<uint64_t>(v);
        note(slot);
    }
    void field_u32(int slot, uint32_t v) {
        if (v == 0) return;
        prep(4, 0);
        push<uint32_t>(v);
        note(slot);
    }
    void field_ref(int slot, Off target) {
        push_ref(target);
        note(slot);
    }
    void field_u8(int slot, uint8_t v) {
        if (v == 0) return;
        push<uint8_t>(v);
        note(slot);
    }
    Off end_table() {
        prep(4, 0);
        push<int32_t>(0);  // soffset to the vtable, patched below
        const Off table = size();
        int max_slot = 2;
        for (int i = 0; i < n_fields_; i++) max_slot = std::max(max_slot, fields_[i].slot);
        const uint16_t vt_len = (uint16_t)(max_slot + 2);
        std::string vt(vt_len, '\0');
        auto put16 = [&](size_t at, uint16_t v) { memcpy(&vt[at], &v, 2); };
        put16(0, vt_len);
        put16(2, (uint16_t)(table - table_end_));
        for (int i = 0; i < n_fields_; i++) put16((size_t)fields_[i].slot, (uint16_t)(table - fields_[i].off));
        Off vt_off;
        auto it = vtables_.find(vt);
        if (it != vtables_.end()) {
            vt_off = it->second;
        } else {
            prep(2, 0);
            push_bytes(vt.data(), vt.size());
            vt_off = size();
            vtables_.emplace(std::move(vt), vt_off);
        }
        // address(table) - address(vtable) = vt_off - table   (offsets count from the end)
        int32_t so = (int32_t)vt_off - (int32_t)table;
        memcpy(&buf_[buf_.size() - table], &so, 4);
        return table;
    }

    // size prefix + root offset + file identifier
    std::vector<uint8_t> finish(Off root, const char ident[4]) {
        prep(minalign_, 4 + 4 + 4);
        push_bytes(ident, 4);
        push_ref(root);
        push<uint32_t>(size());
        return std::vector<uint8_t>(buf_.begin() + head_, buf_.end());
    }

private:
    struct F {
        int slot;
        Off off;
    };
    void note(int slot) {
        fields_[n_fields_].slot = slot;
        fields_[n_fields_].off = size();
        n_fields_++;
    }
    void grow(size_t n) {
        if (head_ >= n) return;
        size_t old = buf_.size(), used = old - head_;
        size_t cap = std::max(old * 2, used + n + 64);
        std::vector<uint8_t> nb(cap);
        memcpy(&nb[cap - used], &buf_[head_], used);
        buf_.swap(nb);
        head_ = cap - used;
    }
    std::vector<uint8_t> buf_;
    size_t head_;
    size_t minalign_ = 4;
    F fields_[8];
    int n_fields_ = 0;
    Off table_end_ = 0;
    std::map<std::string, Off> vtables_;
};

using Off = Builder::Off;

Off w_wire(Builder& b, uint64_t id) {  // wire.rs:23-27
    b.start_table();
    b.field_u64(4, id);
    return b.end_table();
}

Off w_value(Builder& b, const std::vector<uint8_t>& v) {  // value.rs:31-37
    Off data = b.bytes_vector(v.data(), v.size());
    b.start_table();
    b.field_ref(4, data);
    return b.end_table();
}

Off w_wirelist(Builder& b, const WireList& wl) {  // wire.rs:75-143
    std::vector<Off> els(wl.size());
    for (size_t i = 0; i < wl.size(); i++) {
        Off inner;
        if (!wl[i].is_range) {
            inner = w_wire(b, wl[i].first);
        } else {
            Off f = w_wire(b, wl[i].first), l = w_wire(b, wl[i].last);
            b.start_table();
            b.field_ref(4, f);
            b.field_ref(6, l);
            inner = b.end_table();
        }
        b.start_table();
        b.field_ref(6, inner);
        b.field_u8(4, wl[i].is_range ? 2 : 1);
        els[i] = b.end_table();
    }
    Off vec = b.offset_vector(els);
    b.start_table();
    b.field_ref(4, vec);
    return b.end_table();
}

Off w_iterexpr(Builder& b, const IterExpr& e) {  // iterators.rs:112-228
    Off inner;
    switch (e.type) {
        case 1:
            b.start_table();
            b.field_u64(4, e.value);
            inner = b.end_table();
            break;
        case 2: {
            Off name = b.string(e.name);
            b.start_table();
            b.field_ref(4, name);
            inner = b.end_table();
        } break;
        case 3: case 4: case 5: {
            Off l = w_iterexpr(b, *e.l), r = w_iterexpr(b, *e.r);
            b.start_table();
            b.field_ref(4, l);
            b.field_ref(6, r);
            inner = b.end_table();
        } break;
        default: {
            Off n = w_iterexpr(b, *e.l);
            b.start_table();
            b.field_u64(6, e.value);
            b.field_ref(4, n);
            inner = b.end_table();
        }
    }
    b.start_table();
    b.field_ref(6, inner);
    b.field_u8(4, e.type);
    return b.end_table();
}

Off w_iterexpr_list(Builder& b, const IterExprList& l) {  // iterators.rs:271-312
    std::vector<Off> els(l.size());
    for (size_t i = 0; i < l.size(); i++) {
        Off inner;
        if (!l[i].is_range) {
            inner = w_iterexpr(b, l[i].first);
        } else {
            Off f = w_iterexpr(b, l[i].first), la = w_iterexpr(b, l[i].last);
            b.start_table();
            b.field_ref(4, f);
            b.field_ref(6, la);
            inner = b.end_table();
        }
        b.start_table();
        b.field_ref(6, inner);
        b.field_u8(4, l[i].is_range ? 2 : 1);
        els[i] = b.end_table();
    }
    Off vec = b.offset_vector(els);
    b.start_table();
    b.field_ref(4, vec);
    return b.end_table();
}

Off w_gates(Builder& b, const std::vector<Gate>& gates, const std::vector<std::vector<uint8_t>>& consts);

Off w_gate(Builder& b, const Gate& g, const std::vector<std::vector<uint8_t>>& consts) {  // gates.rs:263-679
    Off inner = 0;
    switch (g.type) {
        case G_CONSTANT: {
            const auto& v = consts[g.const_idx];
            Off c = b.bytes_vector(v.data(), v.size());
            Off o = w_wire(b, g.w0);
            b.start_table();
            b.field_ref(4, o);
            b.field_ref(6, c);
            inner = b.end_table();
        } break;
        case G_ASSERT_ZERO: case G_INSTANCE: case G_WITNESS: {
            Off o = w_wire(b, g.w0);
            b.start_table();
            b.field_ref(4, o);
            inner = b.end_table();
        } break;
        case G_COPY: case G_NOT: {
            Off o = w_wire(b, g.w0), i = w_wire(b, g.w1);
            b.start_table();
            b.field_ref(4, o);
            b.field_ref(6, i);
            inner = b.end_table();
        } break;
        case G_ADD: case G_MUL: case G_AND: case G_XOR: {
            Off o = w_wire(b, g.w0), l = w_wire(b, g.w1), r = w_wire(b, g.w2);
            b.start_table();
            b.field_ref(4, o);
            b.field_ref(6, l);
            b.field_ref(8, r);
            inner = b.end_table();
        } break;
        case G_ADD_CONSTANT: case G_MUL_CONSTANT: {
            const auto& v = consts[g.const_idx];
            Off c = b.bytes_vector(v.data(), v.size());
            Off o = w_wire(b, g.w0), i = w_wire(b, g.w1);
            b.start_table();
            b.field_ref(4, o);
            b.field_ref(6, i);
            b.field_ref(8, c);
            inner = b.end_table();
        } break;
        case G_FREE: {
            Off f = w_wire(b, g.w0);
            Off l = g.has_last ? w_wire(b, g.w1) : 0;
            b.start_table();
            b.field_ref(4, f);
            if (g.has_last) b.field_ref(6, l);
            inner = b.end_table();
        } break;
        case G_CALL: {
            Off name = b.string(g.cx->name);
            Off o = w_wirelist(b, g.cx->outputs), i = w_wirelist(b, g.cx->inputs);
            b.start_table();
            b.field_ref(4, name);
            b.field_ref(6, o);
            b.field_ref(8, i);
            inner = b.end_table();
        } break;
        case G_ANON_CALL: {
            Off sub = w_gates(b, g.cx->body, consts);
            Off i = w_wirelist(b, g.cx->inputs);
            b.start_table();
            b.field_u64(6, g.cx->instance_count);
            b.field_u64(8, g.cx->witness_count);
            b.field_ref(4, i);
            b.field_ref(10, sub);
            Off abs = b.end_table();
            Off o = w_wirelist(b, g.cx->outputs);
            b.start_table();
            b.field_ref(4, o);
            b.field_ref(6, abs);
            inner = b.end_table();
        } break;
        case G_SWITCH: {
            std::vector<Off> cases(g.cx->cases.size()), branches(g.cx->branches.size());
            for (size_t k = 0; k < cases.size(); k++) cases[k] = w_value(b, consts[g.cx->cases[k]]);
            for (size_t k = 0; k < branches.size(); k++) {  // function.rs:174-267
                const CaseInvoke& br = g.cx->branches[k];
                Off inv;
                if (!br.is_anon) {
                    Off name = b.string(br.name);
                    Off i = w_wirelist(b, br.inputs);
                    b.start_table();
                    b.field_ref(4, name);
                    b.field_ref(6, i);
                    inv = b.end_table();
                } else {
                    Off sub = w_gates(b, br.subcircuit, consts);
                    Off i = w_wirelist(b, br.inputs);
                    b.start_table();
                    b.field_u64(6, br.instance_count);
                    b.field_u64(8, br.witness_count);
                    b.field_ref(4, i);
                    b.field_ref(10, sub);
                    inv = b.end_table();
                }
                b.start_table();
                b.field_ref(6, inv);
                b.field_u8(4, br.is_anon ? 2 : 1);
                branches[k] = b.end_table();
            }
            Off cv = b.offset_vector(cases), bv = b.offset_vector(branches);
            Off cond = w_wire(b, g.w0);
            Off o = w_wirelist(b, g.cx->outputs);
            b.start_table();
            b.field_ref(4, cond);
            b.field_ref(6, o);
            b.field_ref(8, cv);
            b.field_ref(10, bv);
            inner = b.end_table();
        } break;
        case G_FOR: {
            const Complex& c = *g.cx;
            Off body;
            if (!c.body_is_anon) {
                Off name = b.string(c.fn_name);
                Off o = w_iterexpr_list(b, c.it_outputs), i = w_iterexpr_list(b, c.it_inputs);
                b.start_table();
                b.field_ref(4, name);
                b.field_ref(6, o);
                b.field_ref(8, i);
                body = b.end_table();
            } else {
                Off sub = w_gates(b, c.body, consts);
                Off o = w_iterexpr_list(b, c.it_outputs), i = w_iterexpr_list(b, c.it_inputs);
                b.start_table();
                b.field_u64(8, c.instance_count);
                b.field_u64(10, c.witness_count);
                b.field_ref(4, o);
                b.field_ref(6, i);
                b.field_ref(12, sub);
                body = b.end_table();
            }
            Off it = b.string(c.name);
            Off o = w_wirelist(b, c.outputs);
            b.start_table();
            b.field_u64(8, c.first);
            b.field_u64(10, c.last);
            b.field_ref(4, o);
            b.field_ref(6, it);
            b.field_ref(14, body);
            b.field_u8(12, c.body_is_anon ? 2 : 1);
            inner = b.end_table();
        } break;
        default: break;
    }
    b.start_table();
    b.field_ref(6, inner);
    b.field_u8(4, g.type);
    return b.end_table();
}

Off w_gates(Builder& b, const std::vector<Gate>& gates, const std::vector<std::vector<uint8_t>>& consts) {
    std::vector<Off> els(gates.size());
    // the last gate is built first so that gate 0 ends up at the lowest address, in reading order
    for (size_t i = gates.size(); i-- > 0;) els[i] = w_gate(b, gates[i], consts);
    return b.offset_vector(els);
}

Off w_header(Builder& b, const Header& h) {  // header.rs:58-76
    Off fc = w_value(b, h.field_characteristic);
    Off ver = b.string(h.version);
    b.start_table();
    b.field_ref(4, ver);
    b.field_ref(6, fc);
    b.field_u32(8, h.field_degree);
    return b.end_table();
}

std::string gateset_string(uint16_t m) {  // relation.rs:177-221
    if ((m & M_ARITH) == M_ARITH) return "arithmetic";
    if ((m & M_BOOL) == M_BOOL) return "boolean";
    std::string s;
    if (m & M_ADD) s += "@add,";
    if (m & M_ADDC) s += "@addc,";
    if (m & M_MUL) s += "@mul,";
    if (m & M_MULC) s += "@mulc,";
    if (m & M_XOR) s += "@xor,";
    if (m & M_NOT) s += "@not,";
    if (m & M_AND) s += "@and,";
    return s;
}

std::string feature_string(uint16_t m) {  // relation.rs:258-282
    if ((m & (M_FUNCTION | M_FOR | M_SWITCH)) == 0) return "simple";
    std::string s;
    if (m & M_FOR) s += "@for,";
    if (m & M_SWITCH) s += "@switch,";
    if (m & M_FUNCTION) s += "@function,";
    return s;
}

}  // namespace

// One size-prefixed message (Relation / Instance / Witness ::write_into).
std::vector<uint8_t> write_message(const Message& m) {
    Builder b;
    Off body;
    if (m.type == MSG_RELATION) {
        Off header = w_header(b, m.header);
        Off directives = w_gates(b, m.gates, m.consts);
        std::vector<Off> fns(m.functions.size());
        for (size_t i = m.functions.size(); i-- > 0;) {  // function.rs:48-80
            const Function& f = m.functions[i];
            Off fb = w_gates(b, f.body, m.consts);
            Off name = b.string(f.name);
            b.start_table();
            b.field_u64(6, f.output_count);
            b.field_u64(8, f.input_count);
            b.field_u64(10, f.instance_count);
            b.field_u64(12, f.witness_count);
            b.field_ref(4, name);
            b.field_ref(14, fb);
            fns[i] = b.end_table();
        }
        Off functions = b.offset_vector(fns);
        Off gateset = b.string(gateset_string(m.gate_mask));
        Off features = b.string(feature_string(m.feat_mask));
        b.start_table();
        b.field_ref(4, header);
        b.field_ref(6, gateset);
        b.field_ref(8, features);
        b.field_ref(10, functions);
        b.field_ref(12, directives);
        body = b.end_table();
    } else {
        Off header = w_header(b, m.header);
        std::vector<Off> vals(m.values.size());
        for (size_t i = m.values.size(); i-- > 0;) vals[i] = w_value(b, m.values[i]);
        Off vec = b.offset_vector(vals);
        b.start_table();
        b.field_ref(4, header);
        b.field_ref(6, vec);
        body = b.end_table();
    }
    b.start_table();
    b.field_ref(6, body);
    b.field_u8(4, (uint8_t)m.type);
    Off root = b.end_table();
    return b.finish(root, "siev");
}

}  // namespace ir
}  // namespace zkb
