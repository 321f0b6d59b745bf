// FlatBuffers reader for sieve_ir.fbs — replaces rust/src/sieve_ir_generated.rs (reader half)
// and the TryFrom conversions in rust/src/structs/*.rs.  Every access is bounds-checked.
#include <string.h>

#include <stdexcept>

#include "ir.h"

namespace zkb {
namespace ir {

namespace {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Buf {
    const uint8_t* p;
    size_t n;
    void need(size_t pos, size_t k) const {
        if (pos > n || k > n - pos) throw ParseError("malformed FlatBuffers message (offset out of bounds)");
    }
    uint16_t u16(size_t pos) const { need(pos, 2); uint16_t v; memcpy(&v, p + pos, 2); return v; }
    uint32_t u32(size_t pos) const { need(pos, 4); uint32_t v; memcpy(&v, p + pos, 4); return v; }
    int32_t i32(size_t pos) const { need(pos, 4); int32_t v; memcpy(&v, p + pos, 4); return v; }
    uint64_t u64(size_t pos) const { need(pos, 8); uint64_t v; memcpy(&v, p + pos, 8); return v; }
    uint8_t u8(size_t pos) const { need(pos, 1); return p[pos]; }
    // FlatBuffers offsets may point many vector entries at ONE shared sub-table, which may again hold a vector of shared
    // entries: a few KB then decode to width^depth owned structs.  Every table view is charged against a budget tied to
    // the message length (a well-formed gate costs ~5 views per ~45 bytes; the budget allows 8 per byte).
    mutable uint64_t views = 0;
    void charge() const {
        if (++views > 8 * (uint64_t)n + 4096)
            throw ParseError("zkb: message decodes to more tables than its size allows (shared sub-tables)");
    }
};

// A table view.  Field slot k of the vtable holds the offset of the field inside the table
// (0 / beyond vtable length: absent -> scalar default 0, sub-table None).
struct Table {
    const Buf* b = nullptr;
    size_t pos = 0, vt = 0;
    uint16_t vtlen = 0;
    bool ok = false;

    Table() {}
    Table(const Buf* buf, size_t at) : b(buf), pos(at), ok(true) {
        b->charge();
        int64_t v = (int64_t)at - (int64_t)b->i32(at);
        if (v < 0 || (size_t)v > b->n) throw ParseError("malformed FlatBuffers message (vtable out of bounds)");
        vt = (size_t)v;
        vtlen = b->u16(vt);
    }
    explicit operator bool() const { return ok; }
    size_t off(int slot) const {
        if (slot + 2 > vtlen) return 0;
        return b->u16(vt + slot);
    }
    uint8_t u8(int slot) const { size_t o = off(slot); return o ? b->u8(pos + o) : 0; }
    uint32_t u32(int slot) const { size_t o = off(slot); return o ? b->u32(pos + o) : 0; }
    uint64_t u64(int slot) const { size_t o = off(slot); return o ? b->u64(pos + o) : 0; }
    bool has(int slot) const { return off(slot) != 0; }
    size_t indirect(int slot) const {
        size_t o = off(slot);
        size_t at = pos + o;
        return at + b->u32(at);
    }
    Table table(int slot) const { return has(slot) ? Table(b, indirect(slot)) : Table(); }
    bool string(int slot, std::string& out) const {
        if (!has(slot)) return false;
        size_t at = indirect(slot);
        uint32_t len = b->u32(at);
        b->need(at + 4, len);
        out.assign((const char*)b->p + at + 4, len);
        return true;
    }
    bool bytes(int slot, std::vector<uint8_t>& out) const {
        if (!has(slot)) return false;
        size_t at = indirect(slot);
        uint32_t len = b->u32(at);
        b->need(at + 4, len);
        out.assign(b->p + at + 4, b->p + at + 4 + len);
        return true;
    }
    // vector of tables: returns element count, element i via vec_at
    bool vec(int slot, size_t& at, uint32_t& len) const {
        if (!has(slot)) return false;
        at = indirect(slot);
        len = b->u32(at);
        b->need(at + 4, (size_t)len * 4);
        return true;
    }
    Table vec_at(size_t at, uint32_t i) const {
        size_t e = at + 4 + (size_t)i * 4;
        return Table(b, e + b->u32(e));
    }
};

#define REQ(cond, msg) do { if (!(cond)) throw ParseError(msg); } while (0)

uint64_t wire_id(const Table& t, const char* missing) {
    REQ(t, missing);
    return t.u64(4);  // Wire.id, absent => 0
}

void read_wirelist(const Table& t, WireList& out) {  // structs/wire.rs:145-157
    size_t at; uint32_t n;
    REQ(t.vec(4, at, n), "Missing wire list elements");
    out.reserve(n);
    for (uint32_t i = 0; i < n; i++) {
        Table el = t.vec_at(at, i);
        uint8_t ty = el.u8(4);
        if (ty == 1) {
            Table w = el.table(6);
            REQ(w, "Missing wire");
            out.push_back(WireEl{w.u64(4), 0, false});
        } else if (ty == 2) {
            Table r = el.table(6);
            REQ(r, "Missing range");
            uint64_t f = wire_id(r.table(4), "Missing start value in range");
            uint64_t l = wire_id(r.table(6), "Missing end value in range");
            out.push_back(WireEl{f, l, true});
        } else {
            throw ParseError("Unknown type in WireListElement");
        }
    }
}

void read_iterexpr(const Table& t, IterExpr& e, int depth = 0) {  // structs/iterators.rs:36-110
    REQ(depth < 512, "zkb: iterator expression nested more than 512 levels deep");
    uint8_t ty = t.u8(4);
    REQ(ty >= 1 && ty <= 6, "Unknown Iterator Expression type");
    Table v = t.table(6);
    REQ(v, "Missing iterator expression value");
    e.type = ty;
    switch (ty) {
        case 1: e.value = v.u64(4); break;
        case 2: REQ(v.string(4, e.name), "IterExpr: No name given"); break;
        case 3: case 4: case 5: {
            Table l = v.table(4), r = v.table(6);
            REQ(l, "Missing left operand");
            REQ(r, "Missing right operand");
            e.l.reset(new IterExpr());
            e.r.reset(new IterExpr());
            read_iterexpr(l, *e.l, depth + 1);
            read_iterexpr(r, *e.r, depth + 1);
        } break;
        default: {
            Table nmr = v.table(4);
            REQ(nmr, "Missing numerator");
            e.l.reset(new IterExpr());
            read_iterexpr(nmr, *e.l, depth + 1);
            e.value = v.u64(6);
        }
    }
}

void read_iterexpr_list(const Table& t, IterExprList& out) {  // structs/iterators.rs:314-326
    size_t at; uint32_t n;
    REQ(t.vec(4, at, n), "Missing iterexpr elements");
    out.resize(n);
    for (uint32_t i = 0; i < n; i++) {
        Table el = t.vec_at(at, i);
        uint8_t ty = el.u8(4);
        Table inner = el.table(6);
        if (ty == 1) {
            REQ(inner, "Missing element");
            out[i].is_range = false;
            read_iterexpr(inner, out[i].first);
        } else if (ty == 2) {
            REQ(inner, "Missing element");
            Table f = inner.table(4), l = inner.table(6);
            REQ(f, "Missing first value of range");
            REQ(l, "Missing last value of range");
            out[i].is_range = true;
            read_iterexpr(f, out[i].first);
            read_iterexpr(l, out[i].last);
        } else {
            throw ParseError("Unknown type in IterExprWireListElement");
        }
    }
}

struct Ctx {
    Message* msg;
    int depth = 0;  // nesting of gate vectors (AnonCall / Switch / For bodies): bounded, a hostile message must not
                    // exhaust the stack (the walkers above the reader recurse over the same structure)
    uint32_t add_const(std::vector<uint8_t>&& v) {
        msg->consts.push_back(std::move(v));
        return (uint32_t)msg->consts.size() - 1;
    }
};

void read_gates(Ctx& cx, const Table& parent, int slot, const char* missing, std::vector<Gate>& out);

void read_gate(Ctx& cx, const Table& d, Gate& g) {  // structs/gates.rs:60-259
    uint8_t ty = d.u8(4);
    REQ(ty >= 1 && ty <= 17, "No gate type");
    Table t = d.table(6);
    REQ(t, "Missing directive");
    g.type = ty;
    std::vector<uint8_t> bytes;
    switch (ty) {
        case G_CONSTANT:
            g.w0 = wire_id(t.table(4), "Missing output");
            REQ(t.bytes(6, bytes), "Missing constant");
            g.const_idx = cx.add_const(std::move(bytes));
            break;
        case G_ASSERT_ZERO:
            g.w0 = wire_id(t.table(4), "Missing input");
            break;
        case G_COPY: case G_NOT:
            g.w0 = wire_id(t.table(4), "Missing output");
            g.w1 = wire_id(t.table(6), "Missing input");
            break;
        case G_ADD: case G_MUL: case G_AND: case G_XOR:
            g.w0 = wire_id(t.table(4), "Missing output");
            g.w1 = wire_id(t.table(6), "Missing left input");
            g.w2 = wire_id(t.table(8), "Missing right input");
            break;
        case G_ADD_CONSTANT: case G_MUL_CONSTANT:
            g.w0 = wire_id(t.table(4), "Missing output");
            g.w1 = wire_id(t.table(6), "Missing input");
            REQ(t.bytes(8, bytes), "Missing constant");
            g.const_idx = cx.add_const(std::move(bytes));
            break;
        case G_INSTANCE: case G_WITNESS:
            g.w0 = wire_id(t.table(4), "Missing output");
            break;
        case G_FREE: {
            g.w0 = wire_id(t.table(4), "Missing first wire");
            Table l = t.table(6);
            g.has_last = (bool)l;
            g.w1 = l ? l.u64(4) : 0;
        } break;
        case G_CALL: {
            g.cx = std::make_shared<Complex>();
            REQ(t.string(4, g.cx->name), "Missing function name.");
            Table o = t.table(6), i = t.table(8);
            REQ(o, "Missing outputs");
            REQ(i, "Missing inputs");
            read_wirelist(o, g.cx->outputs);
            read_wirelist(i, g.cx->inputs);
        } break;
        case G_ANON_CALL: {
            g.cx = std::make_shared<Complex>();
            Table inner = t.table(6);
            REQ(inner, "Missing inner AbstractAnonCall");
            Table o = t.table(4), i = inner.table(4);
            REQ(o, "Missing output wires");
            REQ(i, "Missing input wires");
            read_wirelist(o, g.cx->outputs);
            read_wirelist(i, g.cx->inputs);
            g.cx->instance_count = inner.u64(6);
            g.cx->witness_count = inner.u64(8);
            read_gates(cx, inner, 10, "Missing subcircuit", g.cx->body);
        } break;
        case G_SWITCH: {
            g.cx = std::make_shared<Complex>();
            size_t at; uint32_t n;
            REQ(t.vec(8, at, n), "Missing cases values");
            for (uint32_t k = 0; k < n; k++) {
                Table v = t.vec_at(at, k);
                std::vector<uint8_t> val;
                REQ(v.bytes(4, val), "Missing value");
                g.cx->cases.push_back(cx.add_const(std::move(val)));
            }
            g.w0 = wire_id(t.table(4), "Missing condition wire.");
            Table o = t.table(6);
            REQ(o, "Missing output wires");
            read_wirelist(o, g.cx->outputs);
            REQ(t.vec(10, at, n), "Missing branches");
            g.cx->branches.resize(n);
            for (uint32_t k = 0; k < n; k++) {  // structs/function.rs:132-172
                Table ci = t.vec_at(at, k);
                CaseInvoke& br = g.cx->branches[k];
                uint8_t it = ci.u8(4);
                Table inv = ci.table(6);
                if (it == 1) {
                    REQ(inv, "Missing invocation");
                    br.is_anon = false;
                    REQ(inv.string(4, br.name), "Missing function name.");
                    Table iw = inv.table(6);
                    REQ(iw, "Missing inputs");
                    read_wirelist(iw, br.inputs);
                } else if (it == 2) {
                    REQ(inv, "Missing invocation");
                    br.is_anon = true;
                    read_gates(cx, inv, 10, "Missing implementation", br.subcircuit);
                    Table iw = inv.table(4);
                    REQ(iw, "Missing inputs");
                    read_wirelist(iw, br.inputs);
                    br.instance_count = inv.u64(6);
                    br.witness_count = inv.u64(8);
                } else {
                    throw ParseError("No directive type");
                }
            }
        } break;
        case G_FOR: {
            g.cx = std::make_shared<Complex>();
            Table o = t.table(4);
            REQ(o, "missing output list");
            read_wirelist(o, g.cx->outputs);
            uint8_t bt = t.u8(12);
            Table b = t.table(14);
            if (bt == 1) {
                REQ(b, "Missing body");
                g.cx->body_is_anon = false;
                REQ(b.string(4, g.cx->fn_name), "Missing function in function name");
                Table bo = b.table(6), bi = b.table(8);
                REQ(bo, "missing output list");
                REQ(bi, "missing input list");
                read_iterexpr_list(bo, g.cx->it_outputs);
                read_iterexpr_list(bi, g.cx->it_inputs);
            } else if (bt == 2) {
                REQ(b, "Missing body");
                g.cx->body_is_anon = true;
                Table bo = b.table(4), bi = b.table(6);
                REQ(bo, "missing output list");
                REQ(bi, "missing input list");
                read_iterexpr_list(bo, g.cx->it_outputs);
                read_iterexpr_list(bi, g.cx->it_inputs);
                g.cx->instance_count = b.u64(8);
                g.cx->witness_count = b.u64(10);
                read_gates(cx, b, 12, "Missing body", g.cx->body);
            } else {
                throw ParseError("Unknown body type");
            }
            REQ(t.string(6, g.cx->name), "Missing iterator name");
            g.cx->first = t.u64(8);
            g.cx->last = t.u64(10);
        } break;
    }
}

void read_gates(Ctx& cx, const Table& parent, int slot, const char* missing, std::vector<Gate>& out) {
    size_t at; uint32_t n;
    REQ(parent.vec(slot, at, n), missing);
    REQ(cx.depth < 512, "zkb: gates nested more than 512 levels deep");
    cx.depth++;
    out.resize(n);
    for (uint32_t i = 0; i < n; i++) read_gate(cx, parent.vec_at(at, i), out[i]);
    cx.depth--;
}

void read_header(const Table& t, Header& h) {  // structs/header.rs:37-56
    REQ(t, "Missing header");
    REQ(t.string(4, h.version), "Missing version");
    Table fc = t.table(6);
    REQ(fc, "Missing field characteristic");
    REQ(fc.bytes(4, h.field_characteristic), "Missing value");
    h.field_degree = t.u32(8);
}

void read_values(const Table& t, int slot, const char* missing, std::vector<std::vector<uint8_t>>& out) {
    size_t at; uint32_t n;
    REQ(t.vec(slot, at, n), missing);
    out.resize(n);
    for (uint32_t i = 0; i < n; i++) REQ(t.vec_at(at, i).bytes(4, out[i]), "Missing value");
}

std::string strip_spaces(const std::string& s) {
    std::string r;
    for (char c : s)
        if (c != ' ') r.push_back(c);
    return r;
}

template <class F>
void split_commas(const std::string& s, F f) {
    size_t start = 0;
    while (true) {
        size_t c = s.find(',', start);
        f(s.substr(start, c == std::string::npos ? std::string::npos : c - start));
        if (c == std::string::npos) break;
        start = c + 1;
    }
}

}  // namespace

bool parse_gate_set(const std::string& s, uint16_t& mask, std::string& err) {
    uint16_t ret = 0;
    bool done = false, bad = false;
    split_commas(s, [&](const std::string& part) {
        if (done || bad) return;
        std::string t = strip_spaces(part);
        if (t == "arithmetic") { ret = M_ARITH; done = true; }
        else if (t == "boolean") { ret = M_BOOL; done = true; }
        else if (t == "@add") ret |= M_ADD;
        else if (t == "@addc") ret |= M_ADDC;
        else if (t == "@mul") ret |= M_MUL;
        else if (t == "@mulc") ret |= M_MULC;
        else if (t == "@xor") ret |= M_XOR;
        else if (t == "@not") ret |= M_NOT;
        else if (t == "@and") ret |= M_AND;
        else if (t.empty()) {}
        else bad = true;
    });
    if (bad) {
        err = "Unable to parse the following gateset: " + s;
        return false;
    }
    mask = ret;
    return true;
}

bool parse_feature_toggle(const std::string& s, uint16_t& mask, std::string& err) {
    uint16_t ret = 0;
    bool done = false, bad = false;
    std::string bad_part;
    split_commas(s, [&](const std::string& part) {
        if (done || bad) return;
        std::string t = strip_spaces(part);
        if (t == "@function") ret |= M_FUNCTION;
        else if (t == "@for") ret |= M_FOR;
        else if (t == "@switch") ret |= M_SWITCH;
        else if (t == "simple") { ret = M_SIMPLE; done = true; }
        else if (t.empty()) {}
        else { bad = true; bad_part = part; }
    });
    if (bad) {
        err = "Unable to parse following feature toggles " + bad_part;
        return false;
    }
    mask = ret;
    return true;
}

void split_messages(const uint8_t* buf, size_t len, std::vector<std::pair<size_t, size_t>>& out) {
    size_t pos = 0;
    while (pos + 4 <= len) {
        uint32_t size;
        memcpy(&size, buf + pos, 4);
        if (size == 0) break;               // explicit end marker
        if ((size_t)size > len - pos - 4) break;  // truncated: read_exact fails in the reference
        out.push_back({pos, (size_t)size + 4});
        pos += (size_t)size + 4;
    }
}

// Flat view of a relation message: the same tables read_message decodes, visited in place (structs/gates.rs:60-259 without
// the owned `Gate` values).  Every access is bounds-checked like in read_message.
int walk_flat_relation(const uint8_t* buf, size_t len, FlatRelationHead& head, FlatGateFn fn, void* ctx, bool wires, std::string& err) {
    try {
        if (len < 12) return FLAT_OTHER;
        Buf b{buf, len};
        Table root(&b, 4 + (size_t)b.u32(4));
        if (root.u8(4) != MSG_RELATION) return FLAT_OTHER;
        Table m = root.table(6);
        if (!m || !m.has(12)) return FLAT_OTHER;      // the owned-struct reader produces the reference's error text
        size_t at; uint32_t n;
        if (m.vec(10, at, n) && n != 0) return FLAT_OTHER;  // functions
        read_header(m.table(4), head.header);
        std::string gs, ft;
        if (!m.string(6, gs) || !m.string(8, ft)) return FLAT_OTHER;
        if (!parse_gate_set(gs, head.gate_mask, err) || !parse_feature_toggle(ft, head.feat_mask, err)) return FLAT_OTHER;
        if (!m.vec(12, at, n)) return FLAT_OTHER;
        head.n_gates = n;
        FlatGate g;
        for (uint32_t i = 0; i < n; i++) {
            Table d = m.vec_at(at, i);
            const uint8_t ty = d.u8(4);
            if (ty < G_CONSTANT || ty > G_WITNESS || ty == G_COPY) return FLAT_OTHER;  // Copy aliases, Free re-uses ids, the rest is structured
            Table t = d.table(6);
            if (!t) return FLAT_OTHER;
            g.type = ty;
            g.cbytes = nullptr;
            g.clen = 0;
            g.w1 = g.w2 = 0;
            Table w = t.table(4);
            if (!w) return FLAT_OTHER;
            g.w0 = w.u64(4);
            if (ty == G_CONSTANT || ty == G_ADD_CONSTANT || ty == G_MUL_CONSTANT) {
                const int slot = ty == G_CONSTANT ? 6 : 8;
                if (!t.has(slot)) return FLAT_OTHER;
                size_t cat = t.indirect(slot);
                uint32_t clen = b.u32(cat);
                b.need(cat + 4, clen);
                g.cbytes = b.p + cat + 4;
                g.clen = clen;
            }
            if (wires && ty != G_CONSTANT && ty != G_ASSERT_ZERO && ty != G_INSTANCE && ty != G_WITNESS) {
                Table w1 = t.table(6);
                if (!w1) return FLAT_OTHER;
                g.w1 = w1.u64(4);
                if (ty == G_ADD || ty == G_MUL || ty == G_AND || ty == G_XOR) {
                    Table w2 = t.table(8);
                    if (!w2) return FLAT_OTHER;
                    g.w2 = w2.u64(4);
                }
            }
            if (!fn(ctx, g)) return FLAT_STOPPED;
        }
        return FLAT_OK;
    } catch (const ParseError&) {
        return FLAT_OTHER;  // malformed: read_message reports it with the reference's wording when the message is reached
    }
}

bool read_message(const uint8_t* buf, size_t len, Message& out, std::string& err) {
    try {
        REQ(len >= 12, "malformed FlatBuffers message (too short)");
        Buf b{buf, len};
        Table root(&b, 4 + (size_t)b.u32(4));
        uint8_t mt = root.u8(4);
        Table m = root.table(6);
        Ctx cx{&out};
        if (mt == MSG_INSTANCE || mt == MSG_WITNESS) {
            REQ(m, "Missing message");
            out.type = (MsgType)mt;
            read_header(m.table(4), out.header);
            read_values(m, 6, mt == MSG_INSTANCE ? "Missing common_input" : "Missing short_witness", out.values);
        } else if (mt == MSG_RELATION) {
            REQ(m, "Missing message");
            out.type = MSG_RELATION;
            // relation.rs:47-72: directives first, then functions, header, gateset, features, gates
            REQ(m.has(12), "Missing directives");
            size_t at; uint32_t n;
            if (m.vec(10, at, n)) {
                out.functions.resize(n);
                for (uint32_t i = 0; i < n; i++) {  // function.rs:29-46
                    Table f = m.vec_at(at, i);
                    Function& fn = out.functions[i];
                    REQ(f.has(14), "Missing reference implementation");
                    REQ(f.string(4, fn.name), "Missing name");
                    fn.output_count = f.u64(6);
                    fn.input_count = f.u64(8);
                    fn.instance_count = f.u64(10);
                    fn.witness_count = f.u64(12);
                    read_gates(cx, f, 14, "Missing reference implementation", fn.body);
                }
            }
            read_header(m.table(4), out.header);
            std::string gs, ft;
            REQ(m.string(6, gs), "Missing gateset description");
            if (!parse_gate_set(gs, out.gate_mask, err)) return false;
            REQ(m.string(8, ft), "Missing feature toggles");
            if (!parse_feature_toggle(ft, out.feat_mask, err)) return false;
            read_gates(cx, m, 12, "Missing directives", out.gates);
        } else {
            throw ParseError("Invalid message type");
        }
        return true;
    } catch (const ParseError& e) {
        err = e.what();
        return false;
    }
}

}  // namespace ir
}  // namespace zkb
