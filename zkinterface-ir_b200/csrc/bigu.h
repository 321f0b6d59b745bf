// Minimal host-side unsigned big integer: just enough to parse `Value`s
// (little-endian byte strings of any length, rust/src/structs/value.rs:11),
// reduce them modulo p, and derive the Montgomery constants.  Not a hot path.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

namespace zkb {

class BigU {
public:
    std::vector<uint32_t> w;  // little-endian limbs, no trailing zeros

    BigU() {}
    explicit BigU(uint64_t v) {
        if (v) w.push_back((uint32_t)v);
        if (v >> 32) w.push_back((uint32_t)(v >> 32));
    }
    static BigU from_bytes_le(const uint8_t* b, size_t n) {
        BigU r;
        r.w.assign((n + 3) / 4, 0);
        for (size_t i = 0; i < n; i++) r.w[i / 4] |= (uint32_t)b[i] << (8 * (i % 4));
        r.trim();
        return r;
    }
    void trim() {
        while (!w.empty() && w.back() == 0) w.pop_back();
    }
    bool is_zero() const { return w.empty(); }
    bool is_one() const { return w.size() == 1 && w[0] == 1; }
    size_t bits() const {
        if (w.empty()) return 0;
        return 32 * (w.size() - 1) + (32 - __builtin_clz(w.back()));
    }
    bool bit(size_t i) const {
        size_t k = i / 32;
        return k < w.size() && ((w[k] >> (i % 32)) & 1);
    }
    uint32_t limb(size_t i) const { return i < w.size() ? w[i] : 0; }
    static int cmp(const BigU& a, const BigU& b) {
        if (a.w.size() != b.w.size()) return a.w.size() < b.w.size() ? -1 : 1;
        for (size_t i = a.w.size(); i-- > 0;)
            if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
        return 0;
    }
    bool operator<(const BigU& o) const { return cmp(*this, o) < 0; }
    bool operator==(const BigU& o) const { return cmp(*this, o) == 0; }
    bool operator>=(const BigU& o) const { return cmp(*this, o) >= 0; }

    void add(const BigU& o) {
        size_t n = std::max(w.size(), o.w.size());
        w.resize(n, 0);
        uint64_t c = 0;
        for (size_t i = 0; i < n; i++) {
            c += (uint64_t)w[i] + o.limb(i);
            w[i] = (uint32_t)c;
            c >>= 32;
        }
        if (c) w.push_back((uint32_t)c);
    }
    // requires *this >= o
    void sub(const BigU& o) {
        uint64_t br = 0;
        for (size_t i = 0; i < w.size(); i++) {
            uint64_t t = (uint64_t)w[i] - o.limb(i) - br;
            w[i] = (uint32_t)t;
            br = (t >> 63) & 1;
        }
        trim();
    }
    void shl1() {
        uint32_t c = 0;
        for (size_t i = 0; i < w.size(); i++) {
            uint32_t n = w[i] >> 31;
            w[i] = (w[i] << 1) | c;
            c = n;
        }
        if (c) w.push_back(c);
    }
    void shr1() {
        uint32_t c = 0;
        for (size_t i = w.size(); i-- > 0;) {
            uint32_t n = w[i] & 1;
            w[i] = (w[i] >> 1) | (c << 31);
            c = n;
        }
        trim();
    }
    // *this mod m, by binary long division (m != 0)
    BigU mod(const BigU& m) const {
        if (*this < m) return *this;
        BigU r;
        for (size_t i = bits(); i-- > 0;) {
            r.shl1();
            if (bit(i)) {
                if (r.w.empty()) r.w.push_back(1);
                else r.w[0] |= 1;
            }
            if (r >= m) r.sub(m);
        }
        return r;
    }
    // (a*b) mod m by double-and-add (a, b < m)
    static BigU mulmod(const BigU& a, const BigU& b, const BigU& m) {
        BigU r;
        for (size_t i = b.bits(); i-- > 0;) {
            r.shl1();
            if (r >= m) r.sub(m);
            if (b.bit(i)) {
                r.add(a);
                if (r >= m) r.sub(m);
            }
        }
        return r;
    }
    void to_limbs(uint32_t* out, int n) const {
        for (int i = 0; i < n; i++) out[i] = limb(i);
    }
    static BigU from_limbs(const uint32_t* in, int n) {
        BigU r;
        r.w.assign(in, in + n);
        r.trim();
        return r;
    }
    std::string to_dec() const {
        if (w.empty()) return "0";
        std::vector<uint32_t> t = w;
        std::string s;
        while (!t.empty()) {
            uint64_t rem = 0;
            for (size_t i = t.size(); i-- > 0;) {
                uint64_t cur = (rem << 32) | t[i];
                t[i] = (uint32_t)(cur / 10);
                rem = cur % 10;
            }
            s.push_back((char)('0' + rem));
            while (!t.empty() && t.back() == 0) t.pop_back();
        }
        std::reverse(s.begin(), s.end());
        return s;
    }
};

}  // namespace zkb
