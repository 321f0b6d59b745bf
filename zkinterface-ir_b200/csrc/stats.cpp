// zkb_metrics — mirror of `Stats` (rust/src/consumers/stats.rs:11-287): the gate / message counts the `metrics` and
// `valid-eval-metrics` verbs print as JSON (cli.rs:322-363).  Host only.
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/zkb.h"
#include "file_bytes.h"
#include "ir.h"

using namespace zkb;

namespace zkb {
int list_workspace_files(const char* const* paths, size_t n, std::vector<std::string>& out, std::string& err);
}

namespace {

enum Field {  // GateStats, stats.rs:11-41, declaration order (= serde's field order)
    INSTANCE_VARIABLES, WITNESS_VARIABLES, CONSTANTS_GATES, ASSERT_ZERO_GATES, COPY_GATES, ADD_GATES, MUL_GATES, ADD_CONSTANT_GATES,
    MUL_CONSTANT_GATES, AND_GATES, XOR_GATES, NOT_GATES, VARIABLES_FREED, FUNCTIONS_DEFINED, FUNCTIONS_CALLED, SWITCHES, BRANCHES,
    FOR_LOOPS, INSTANCE_MESSAGES, WITNESS_MESSAGES, RELATION_MESSAGES, N_FIELDS
};
const char* const kNames[N_FIELDS] = {"instance_variables", "witness_variables", "constants_gates", "assert_zero_gates", "copy_gates",
                                      "add_gates", "mul_gates", "add_constant_gates", "mul_constant_gates", "and_gates", "xor_gates",
                                      "not_gates", "variables_freed", "functions_defined", "functions_called", "switches", "branches",
                                      "for_loops", "instance_messages", "witness_messages", "relation_messages"};

struct GateStats {
    uint64_t v[N_FIELDS] = {0};
    // ingest_call_stats, stats.rs:268-286: gates, frees, switches/branches/loops and calls; NOT the variables or messages
    void add_call(const GateStats& o) {
        for (int f = CONSTANTS_GATES; f <= VARIABLES_FREED; f++) v[f] += o.v[f];
        v[SWITCHES] += o.v[SWITCHES];
        v[BRANCHES] += o.v[BRANCHES];
        v[FOR_LOOPS] += o.v[FOR_LOOPS];
        v[FUNCTIONS_CALLED] += o.v[FUNCTIONS_CALLED];
    }
};

struct FnStats {
    GateStats stats;
    uint64_t instance_count = 0, witness_count = 0;
};
using Known = std::map<std::string, FnStats>;

void ingest_gate(GateStats& st, const ir::Gate& g, const Known& known);

GateStats ingest_subcircuit(const std::vector<ir::Gate>& sub, const Known& known) {  // stats.rs:114-123
    GateStats local;
    for (const auto& g : sub) ingest_gate(local, g, known);
    return local;
}

// a named invocation: counts the call, folds the callee's stats in; false when the function is unknown
bool named_call(GateStats& st, const std::string& name, const Known& known, uint64_t& ic, uint64_t& wc) {
    st.v[FUNCTIONS_CALLED]++;
    ic = wc = 0;
    auto it = known.find(name);
    if (it == known.end()) {
        fprintf(stderr, "WARNING Stats: function not defined \"%s\"\n", name.c_str());
        return false;
    }
    st.add_call(it->second.stats);
    ic = it->second.instance_count;
    wc = it->second.witness_count;
    return true;
}

void ingest_gate(GateStats& st, const ir::Gate& g, const Known& known) {  // stats.rs:126-266
    uint64_t ic, wc;
    switch (g.type) {
        case ir::G_CONSTANT: st.v[CONSTANTS_GATES]++; break;
        case ir::G_ASSERT_ZERO: st.v[ASSERT_ZERO_GATES]++; break;
        case ir::G_COPY: st.v[COPY_GATES]++; break;
        case ir::G_ADD: st.v[ADD_GATES]++; break;
        case ir::G_MUL: st.v[MUL_GATES]++; break;
        case ir::G_ADD_CONSTANT: st.v[ADD_CONSTANT_GATES]++; break;
        case ir::G_MUL_CONSTANT: st.v[MUL_CONSTANT_GATES]++; break;
        case ir::G_AND: st.v[AND_GATES]++; break;
        case ir::G_XOR: st.v[XOR_GATES]++; break;
        case ir::G_NOT: st.v[NOT_GATES]++; break;
        case ir::G_INSTANCE: st.v[INSTANCE_VARIABLES]++; break;
        case ir::G_WITNESS: st.v[WITNESS_VARIABLES]++; break;
        case ir::G_FREE: st.v[VARIABLES_FREED] += (g.has_last ? g.w1 : g.w0) - g.w0 + 1; break;  // wraps like the release build
        case ir::G_CALL:
            if (named_call(st, g.cx->name, known, ic, wc)) {
                st.v[INSTANCE_VARIABLES] += ic;
                st.v[WITNESS_VARIABLES] += wc;
            }
            break;
        case ir::G_ANON_CALL:
            st.add_call(ingest_subcircuit(g.cx->body, known));
            st.v[INSTANCE_VARIABLES] += g.cx->instance_count;
            st.v[WITNESS_VARIABLES] += g.cx->witness_count;
            break;
        case ir::G_SWITCH: {
            st.v[SWITCHES]++;
            st.v[BRANCHES] += g.cx->branches.size();
            uint64_t mi = 0, mw = 0;
            for (const auto& br : g.cx->branches) {
                if (!br.is_anon) {
                    named_call(st, br.name, known, ic, wc);
                } else {
                    st.add_call(ingest_subcircuit(br.subcircuit, known));
                    ic = br.instance_count;
                    wc = br.witness_count;
                }
                mi = std::max(mi, ic);
                mw = std::max(mw, wc);
            }
            st.v[INSTANCE_VARIABLES] += mi;
            st.v[WITNESS_VARIABLES] += mw;
        } break;
        case ir::G_FOR: {
            const ir::Complex& cx = *g.cx;
            st.v[FOR_LOOPS]++;
            if (cx.last < cx.first) break;
            // every iteration adds the same amounts: the body's statistics do not depend on the iterator
            const uint64_t iters = cx.last - cx.first + 1;
            GateStats once;
            uint64_t i1 = 0, w1 = 0;
            if (!cx.body_is_anon) {
                if (named_call(once, cx.fn_name, known, ic, wc)) {
                    i1 = ic;
                    w1 = wc;
                }
            } else {
                once.add_call(ingest_subcircuit(cx.body, known));
                i1 = cx.instance_count;
                w1 = cx.witness_count;
            }
            for (int f = 0; f < N_FIELDS; f++) st.v[f] += once.v[f] * iters;
            st.v[INSTANCE_VARIABLES] += i1 * iters;
            st.v[WITNESS_VARIABLES] += w1 * iters;
        } break;
        default: break;
    }
}

void json_gate_stats(std::string& out, const GateStats& s, const std::string& indent) {
    out += "{\n";
    for (int f = 0; f < N_FIELDS; f++) {
        out += indent + "  \"" + kNames[f] + "\": " + std::to_string((unsigned long long)s.v[f]);
        out += f + 1 < N_FIELDS ? ",\n" : "\n";
    }
    out += indent + "}";
}

// JSON string literal; bytes that are not valid UTF-8 (a corrupted message: FlatBuffers strings are not validated by the
// reference's reader either) become U+FFFD so that the output is always valid JSON
std::string json_string(const std::string& s) {
    std::string o = "\"";
    for (size_t i = 0; i < s.size();) {
        unsigned char c = (unsigned char)s[i];
        if (c == '"' || c == '\\') {
            o += '\\';
            o += (char)c;
            i++;
        } else if (c < 0x20) {
            char b[8];
            snprintf(b, sizeof b, "\\u%04x", c);
            o += b;
            i++;
        } else if (c < 0x80) {
            o += (char)c;
            i++;
        } else {
            int len = (c >= 0xC2 && c <= 0xDF) ? 2 : (c >= 0xE0 && c <= 0xEF) ? 3 : (c >= 0xF0 && c <= 0xF4) ? 4 : 0;
            bool ok = len != 0 && i + (size_t)len <= s.size();
            for (int k = 1; ok && k < len; k++) ok = ((unsigned char)s[i + k] & 0xC0) == 0x80;
            if (ok && len == 3) {
                unsigned char c1 = (unsigned char)s[i + 1];
                ok = !(c == 0xE0 && c1 < 0xA0) && !(c == 0xED && c1 >= 0xA0);  // overlong / surrogates
            }
            if (ok && len == 4) {
                unsigned char c1 = (unsigned char)s[i + 1];
                ok = !(c == 0xF0 && c1 < 0x90) && !(c == 0xF4 && c1 >= 0x90);
            }
            if (ok) {
                o.append(s, i, (size_t)len);
                i += (size_t)len;
            } else {
                o += "\\ufffd";
                i++;
            }
        }
    }
    return o + "\"";
}

}  // namespace

struct zkb_metrics {
    std::vector<uint8_t> field_characteristic;  // Stats, stats.rs:43-53
    uint32_t field_degree = 0;
    GateStats gate_stats;
    Known functions;
    std::string err, json;

    int fail(int code, const std::string& m) {
        err = m;
        return code;
    }
    int ingest_bytes(const uint8_t* buf, size_t len) {
        ir::Message m;
        std::string e;
        if (!ir::read_message(buf, len, m, e)) return fail(ZKB_E_FORMAT, e);  // from_messages unwraps (stats.rs:58)
        field_characteristic = m.header.field_characteristic;  // ingest_header: the last header wins
        field_degree = m.header.field_degree;
        if (m.type == ir::MSG_INSTANCE) {
            gate_stats.v[INSTANCE_MESSAGES]++;
        } else if (m.type == ir::MSG_WITNESS) {
            gate_stats.v[WITNESS_MESSAGES]++;
        } else {
            gate_stats.v[RELATION_MESSAGES]++;
            for (const auto& f : m.functions) {
                gate_stats.v[FUNCTIONS_DEFINED]++;
                FnStats fs;
                fs.stats = ingest_subcircuit(f.body, functions);
                fs.instance_count = f.instance_count;
                fs.witness_count = f.witness_count;
                functions[f.name] = fs;
            }
            for (const auto& g : m.gates) ingest_gate(gate_stats, g, functions);
        }
        return ZKB_OK;
    }
    // serde_json::to_writer_pretty(&stats): two-space indentation, fields in declaration order
    const std::string& to_json() {
        json = "{\n  \"field_characteristic\": [";
        for (size_t i = 0; i < field_characteristic.size(); i++) {
            json += i ? ",\n    " : "\n    ";
            json += std::to_string((unsigned)field_characteristic[i]);
        }
        json += field_characteristic.empty() ? "],\n" : "\n  ],\n";
        json += "  \"field_degree\": " + std::to_string(field_degree) + ",\n  \"gate_stats\": ";
        json_gate_stats(json, gate_stats, "  ");
        json += ",\n  \"functions\": {";
        bool first = true;
        for (const auto& kv : functions) {
            json += first ? "\n" : ",\n";
            first = false;
            json += "    " + json_string(kv.first) + ": [\n      ";
            json_gate_stats(json, kv.second.stats, "      ");
            json += ",\n      " + std::to_string((unsigned long long)kv.second.instance_count) + ",\n      " +
                    std::to_string((unsigned long long)kv.second.witness_count) + "\n    ]";
        }
        json += functions.empty() ? "}\n}" : "\n  }\n}";
        return json;
    }
};

extern "C" zkb_metrics* zkb_metrics_create(void) { return new zkb_metrics(); }
extern "C" void zkb_metrics_destroy(zkb_metrics* m) { delete m; }
extern "C" const char* zkb_metrics_last_error(zkb_metrics* m) { return m->err.c_str(); }
extern "C" int zkb_metrics_ingest_message(zkb_metrics* m, const uint8_t* buf, size_t len) { return m->ingest_bytes(buf, len); }
extern "C" int zkb_metrics_ingest_buffer(zkb_metrics* m, const uint8_t* buf, size_t len) {
    std::vector<std::pair<size_t, size_t>> msgs;
    ir::split_messages(buf, len, msgs);
    for (auto& x : msgs) {
        int rc = m->ingest_bytes(buf + x.first, x.second);
        if (rc != ZKB_OK) return rc;
    }
    return ZKB_OK;
}
extern "C" int zkb_metrics_ingest_paths(zkb_metrics* m, const char* const* paths, size_t n_paths) {
    std::vector<std::string> files;
    std::string e;
    int rc = zkb::list_workspace_files(paths, n_paths, files, e);
    if (rc != ZKB_OK) return m->fail(rc, e);
    for (const auto& f : files) {
        zkb::FileBytes data;
        if (!data.open(f)) {
            fprintf(stderr, "Warning: failed to open file %s\n", f.c_str());
            continue;
        }
        rc = zkb_metrics_ingest_buffer(m, data.data, data.size);
        if (rc != ZKB_OK) return rc;
    }
    return ZKB_OK;
}
extern "C" const char* zkb_metrics_json(zkb_metrics* m) { return m->to_json().c_str(); }
