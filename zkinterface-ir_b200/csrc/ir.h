// Owned structures of SIEVE IR messages and their FlatBuffers reader.
//
// Mirrors rust/src/structs/{gates.rs:18-55, wire.rs:11-19, iterators.rs:17-30, function.rs:19-26,
// 121-126, 269-274, header.rs:12-16, relation.rs:35-41, instance.rs:13-16, witness.rs:13-16}.
// The reader is a hand-written vtable walker over `sieve_ir.fbs` (no flatc / flatbuffers
// runtime in this image); vtable slots are the VT_* constants of rust/src/sieve_ir_generated.rs.
#pragma once
#include <stdint.h>

#include <memory>
#include <string>
#include <vector>

namespace zkb {
namespace ir {

// DirectiveSet discriminants, sieve_ir_generated.rs:422-442
enum GateType : uint8_t {
    G_NONE = 0, G_CONSTANT, G_ASSERT_ZERO, G_COPY, G_ADD, G_MUL, G_ADD_CONSTANT, G_MUL_CONSTANT, G_AND, G_XOR, G_NOT,
    G_INSTANCE, G_WITNESS, G_FREE, G_CALL, G_ANON_CALL, G_SWITCH, G_FOR
};

// gate-set / feature masks, structs/relation.rs:15-32
constexpr uint16_t M_ADD = 0x0001, M_ADDC = 0x0002, M_MUL = 0x0004, M_MULC = 0x0008, M_ARITH = 0x000F;
constexpr uint16_t M_XOR = 0x0100, M_AND = 0x0200, M_NOT = 0x0400, M_BOOL = 0x0700;
constexpr uint16_t M_FUNCTION = 0x1000, M_FOR = 0x2000, M_SWITCH = 0x4000, M_SIMPLE = 0;

struct WireEl {  // WireListElement: Wire(first) or WireRange(first, last)
    uint64_t first, last;
    bool is_range;
};
using WireList = std::vector<WireEl>;

// c0 + sum coef_k * iterator(name_k) over Z / 2^64 (the reference's wrapping u64 arithmetic): what an expression made
// of Const / Name / Add / Sub / Mul-by-a-constant-side is, evaluated without walking the tree
struct AffineIterExpr {
    uint64_t c0 = 0;
    std::vector<std::pair<std::string, uint64_t>> terms;  // first-occurrence order (= the tree's evaluation order)
};

struct IterExpr {  // IterExprWireNumber, iterators.rs:17-30
    uint8_t type = 0;  // 1 Const, 2 Name, 3 Add, 4 Sub, 5 Mul, 6 DivConst
    uint64_t value = 0;  // Const value / DivConst denominator
    std::string name;
    std::unique_ptr<IterExpr> l, r;
    // filled by the Evaluator on first use (one evaluator thread per message)
    mutable bool affine_tried = false;
    mutable std::shared_ptr<const AffineIterExpr> affine;
};
struct IterExprEl {  // Single(first) or Range(first, last)
    bool is_range = false;
    IterExpr first, last;
};
using IterExprList = std::vector<IterExprEl>;

struct Gate;
struct Complex;

struct Gate {  // 40 bytes for the simple gates that make up almost all of a large relation
    uint8_t type = G_NONE;
    bool has_last = false;       // Free(first, Some(last))
    uint32_t const_idx = 0;      // Constant / AddConstant / MulConstant: index into Relation-level `consts`
    uint64_t w0 = 0, w1 = 0, w2 = 0;  // (out, left/in, right) | (in) | (first, last) | Switch condition
    std::shared_ptr<Complex> cx;  // Call / AnonCall / Switch / For payload
};

struct CaseInvoke {  // function.rs:121-126
    bool is_anon = false;
    std::string name;
    WireList inputs;
    uint64_t instance_count = 0, witness_count = 0;
    std::vector<Gate> subcircuit;
};

struct Complex {
    std::string name;            // Call: function name; For: iterator name
    WireList outputs, inputs;    // Call / AnonCall / Switch outputs; For: global output list
    uint64_t instance_count = 0, witness_count = 0;
    std::vector<Gate> body;      // AnonCall subcircuit / For anon body
    // Switch
    std::vector<uint32_t> cases;  // const indices
    std::vector<CaseInvoke> branches;
    // For
    uint64_t first = 0, last = 0;
    bool body_is_anon = false;
    std::string fn_name;
    IterExprList it_outputs, it_inputs;
};

struct Function {  // function.rs:19-26
    std::string name;
    uint64_t output_count = 0, input_count = 0, instance_count = 0, witness_count = 0;
    std::vector<Gate> body;
};

struct Header {
    std::string version;
    std::vector<uint8_t> field_characteristic;
    uint32_t field_degree = 0;
};

enum MsgType { MSG_NONE = 0, MSG_RELATION = 1, MSG_INSTANCE = 2, MSG_WITNESS = 3 };  // sieve_ir_generated.rs:23-29

struct Message {
    MsgType type = MSG_NONE;
    Header header;
    // Instance.common_inputs / Witness.short_witness
    std::vector<std::vector<uint8_t>> values;
    // Relation
    uint16_t gate_mask = 0, feat_mask = 0;
    std::vector<Function> functions;
    std::vector<Gate> gates;
    std::vector<std::vector<uint8_t>> consts;  // constant byte strings referenced by Gate::const_idx
};

// Parse ONE size-prefixed message (Message::try_from, structs/message.rs:15-36).
// Returns false and sets err (the reference's "Missing ..." texts) on malformed input.
bool read_message(const uint8_t* buf, size_t len, Message& out, std::string& err);

// A relation message made of simple, value-defining gates only (what a GateBuilder emits for a flat circuit: one message
// per 100 000 gates) can be visited IN PLACE, without owned structs: head = header / gate set / features, fn is called once
// per top-level gate with the wire ids (and, for Constant / AddConstant / MulConstant, a pointer to the constant's bytes
// inside buf).  wires = false skips the input-wire tables (a counting pass).  Returns FLAT_OK, FLAT_STOPPED (fn returned
// false) or FLAT_OTHER: anything else — another message type, functions, a structured gate, Copy, Free, or a malformed
// table — is left to read_message, which also produces the reference's error texts.
struct FlatRelationHead {
    Header header;
    uint16_t gate_mask = 0, feat_mask = 0;
    uint32_t n_gates = 0;
};
struct FlatGate {
    uint8_t type;
    uint64_t w0, w1, w2;
    const uint8_t* cbytes;
    uint32_t clen;
};
typedef bool (*FlatGateFn)(void* ctx, const FlatGate& g);
enum { FLAT_OK = 0, FLAT_STOPPED = 1, FLAT_OTHER = 2 };
int walk_flat_relation(const uint8_t* buf, size_t len, FlatRelationHead& head, FlatGateFn fn, void* ctx, bool wires, std::string& err);

// Size-prefixed framing, consumers/utils.rs:6-41: offsets/lengths of the messages in a stream.
// A message running past the end of the buffer stops the split (the reference's read_exact fails).
void split_messages(const uint8_t* buf, size_t len, std::vector<std::pair<size_t, size_t>>& out);

// One size-prefixed message from owned structs (Relation / Instance / Witness ::write_into, relation.rs:86-137).
std::vector<uint8_t> write_message(const Message& m);

bool parse_gate_set(const std::string& s, uint16_t& mask, std::string& err);        // relation.rs:144-167
bool parse_feature_toggle(const std::string& s, uint16_t& mask, std::string& err);  // relation.rs:229-244

}  // namespace ir
}  // namespace zkb
