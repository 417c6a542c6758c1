// gojson.cpp — see gojson.hpp.  Escaping table follows Go 1.24
// encoding/json appendString(escapeHTML=true); the reference pins go1.24.3
// (go.mod:3-5) and calls json.Marshal at internal/cdi/fti/fm/client.go:144.
#include "gojson.hpp"

#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace cro {
namespace gojson {

namespace {
const char kHex[] = "0123456789abcdef";

// Length of the valid UTF-8 sequence starting at s[i] (Go utf8.DecodeRune
// acceptance), or 0 when Go would return (RuneError, 1).
size_t utf8_len(const std::string& s, size_t i, unsigned* rune) {
    const unsigned char c0 = (unsigned char)s[i];
    const size_t n = s.size() - i;
    auto cont = [&](size_t k, unsigned lo, unsigned hi) {
        const unsigned char c = (unsigned char)s[i + k];
        return c >= lo && c <= hi;
    };
    if (c0 >= 0xC2 && c0 <= 0xDF) {
        if (n < 2 || !cont(1, 0x80, 0xBF)) return 0;
        *rune = ((c0 & 0x1Fu) << 6) | ((unsigned char)s[i + 1] & 0x3Fu);
        return 2;
    }
    if (c0 >= 0xE0 && c0 <= 0xEF) {
        unsigned lo = 0x80, hi = 0xBF;
        if (c0 == 0xE0) lo = 0xA0;
        if (c0 == 0xED) hi = 0x9F;
        if (n < 3 || !cont(1, lo, hi) || !cont(2, 0x80, 0xBF)) return 0;
        *rune = ((c0 & 0x0Fu) << 12) | (((unsigned char)s[i + 1] & 0x3Fu) << 6) |
                ((unsigned char)s[i + 2] & 0x3Fu);
        return 3;
    }
    if (c0 >= 0xF0 && c0 <= 0xF4) {
        unsigned lo = 0x80, hi = 0xBF;
        if (c0 == 0xF0) lo = 0x90;
        if (c0 == 0xF4) hi = 0x8F;
        if (n < 4 || !cont(1, lo, hi) || !cont(2, 0x80, 0xBF) || !cont(3, 0x80, 0xBF)) return 0;
        *rune = ((c0 & 0x07u) << 18) | (((unsigned char)s[i + 1] & 0x3Fu) << 12) |
                (((unsigned char)s[i + 2] & 0x3Fu) << 6) | ((unsigned char)s[i + 3] & 0x3Fu);
        return 4;
    }
    return 0;
}
}  // namespace

void append_string(std::string& out, const std::string& s) {
    out.push_back('"');
    size_t i = 0;
    while (i < s.size()) {
        const unsigned char b = (unsigned char)s[i];
        if (b < 0x80) {
            const bool safe = b >= 0x20 && b != '"' && b != '\\' && b != '<' && b != '>' && b != '&';
            if (safe) {
                out.push_back((char)b);
            } else {
                switch (b) {
                    case '\\': out += "\\\\"; break;
                    case '"': out += "\\\""; break;
                    case '\b': out += "\\b"; break;
                    case '\f': out += "\\f"; break;
                    case '\n': out += "\\n"; break;
                    case '\r': out += "\\r"; break;
                    case '\t': out += "\\t"; break;
                    default:
                        out += "\\u00";
                        out.push_back(kHex[b >> 4]);
                        out.push_back(kHex[b & 0xF]);
                }
            }
            ++i;
            continue;
        }
        unsigned rune = 0;
        const size_t len = utf8_len(s, i, &rune);
        if (len == 0) {
            out += "\\ufffd";
            ++i;
            continue;
        }
        if (rune == 0x2028 || rune == 0x2029) {
            out += "\\u202";
            out.push_back(kHex[rune & 0xF]);
        } else {
            out.append(s, i, len);
        }
        i += len;
    }
    out.push_back('"');
}

void Writer::comma() {
    if (after_key_) {
        after_key_ = false;
        return;
    }
    if (!first_.empty()) {
        if (!first_.back()) out_.push_back(',');
        first_.back() = false;
    }
}
Writer& Writer::begin_object() { comma(); out_.push_back('{'); first_.push_back(true); return *this; }
Writer& Writer::end_object() { out_.push_back('}'); first_.pop_back(); return *this; }
Writer& Writer::begin_array() { comma(); out_.push_back('['); first_.push_back(true); return *this; }
Writer& Writer::end_array() { out_.push_back(']'); first_.pop_back(); return *this; }
Writer& Writer::key(const char* k) {
    comma();
    append_string(out_, k);
    out_.push_back(':');
    after_key_ = true;
    return *this;
}
Writer& Writer::value(const std::string& s) { comma(); append_string(out_, s); return *this; }
Writer& Writer::value(const char* s) { return value(std::string(s ? s : "")); }
Writer& Writer::value(long long v) { comma(); out_ += std::to_string(v); return *this; }
Writer& Writer::value_u64(unsigned long long v) { comma(); out_ += std::to_string(v); return *this; }
Writer& Writer::value(bool v) { comma(); out_ += v ? "true" : "false"; return *this; }
Writer& Writer::null() { comma(); out_ += "null"; return *this; }
Writer& Writer::field_omitempty(const char* k, const std::string& s) {
    if (!s.empty()) key(k).value(s);
    return *this;
}
Writer& Writer::field_omitempty(const char* k, bool v) {
    if (v) key(k).value(v);
    return *this;
}
Writer& Writer::field_omitempty(const char* k, long long v) {
    if (v != 0) key(k).value(v);
    return *this;
}
Writer& Writer::string_map(const std::map<std::string, std::string>& m) {
    begin_object();
    for (const auto& kv : m) key(kv.first.c_str()).value(kv.second);  // std::map: bytewise order
    return end_object();
}
Writer& Writer::raw(const std::string& json) { comma(); out_ += json; return *this; }

// ---------------------------------------------------------------------------
// reader
// ---------------------------------------------------------------------------
const Value* Value::get(const std::string& k) const {
    const Value* found = nullptr;
    for (const auto& kv : obj)
        if (kv.first == k) found = kv.second.get();
    return found;
}
std::string Value::get_string(const std::string& k, const std::string& dflt) const {
    const Value* v = get(k);
    return (v && v->kind == String) ? v->str : dflt;
}
bool Value::get_bool(const std::string& k, bool dflt) const {
    const Value* v = get(k);
    return (v && v->kind == Bool) ? v->b : dflt;
}
long long Value::get_int(const std::string& k, long long dflt) const {
    const Value* v = get(k);
    return (v && v->kind == Number) ? (v->is_int ? v->inum : (long long)v->num) : dflt;
}

namespace {
struct Parser {
    const std::string& t;
    size_t i = 0;
    std::string err;
    int depth = 0;
    explicit Parser(const std::string& s) : t(s) {}

    void ws() {
        while (i < t.size() && (t[i] == ' ' || t[i] == '\t' || t[i] == '\n' || t[i] == '\r')) ++i;
    }
    bool fail(const std::string& m) {
        if (err.empty()) err = m + " at offset " + std::to_string(i);
        return false;
    }
    static void put_utf8(std::string& o, unsigned r) {
        if (r < 0x80) o.push_back((char)r);
        else if (r < 0x800) { o.push_back((char)(0xC0 | (r >> 6))); o.push_back((char)(0x80 | (r & 0x3F))); }
        else if (r < 0x10000) {
            o.push_back((char)(0xE0 | (r >> 12))); o.push_back((char)(0x80 | ((r >> 6) & 0x3F)));
            o.push_back((char)(0x80 | (r & 0x3F)));
        } else {
            o.push_back((char)(0xF0 | (r >> 18))); o.push_back((char)(0x80 | ((r >> 12) & 0x3F)));
            o.push_back((char)(0x80 | ((r >> 6) & 0x3F))); o.push_back((char)(0x80 | (r & 0x3F)));
        }
    }
    bool hex4(unsigned* r) {
        if (i + 4 > t.size()) return fail("short \\u escape");
        unsigned v = 0;
        for (int k = 0; k < 4; ++k) {
            const char c = t[i + k];
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (unsigned)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (unsigned)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (unsigned)(c - 'A' + 10);
            else return fail("bad \\u escape");
        }
        i += 4;
        *r = v;
        return true;
    }
    bool string(std::string* o) {
        if (i >= t.size() || t[i] != '"') return fail("expected string");
        ++i;
        while (i < t.size()) {
            const unsigned char c = (unsigned char)t[i];
            if (c == '"') { ++i; return true; }
            if (c < 0x20) return fail("control character in string");
            if (c != '\\') { o->push_back((char)c); ++i; continue; }
            if (++i >= t.size()) break;
            const char e = t[i++];
            switch (e) {
                case '"': o->push_back('"'); break;
                case '\\': o->push_back('\\'); break;
                case '/': o->push_back('/'); break;
                case 'b': o->push_back('\b'); break;
                case 'f': o->push_back('\f'); break;
                case 'n': o->push_back('\n'); break;
                case 'r': o->push_back('\r'); break;
                case 't': o->push_back('\t'); break;
                case 'u': {
                    unsigned r;
                    if (!hex4(&r)) return false;
                    if (r >= 0xD800 && r <= 0xDBFF && i + 6 <= t.size() && t[i] == '\\' && t[i + 1] == 'u') {
                        const size_t save = i;
                        i += 2;
                        unsigned r2;
                        if (!hex4(&r2)) return false;
                        if (r2 >= 0xDC00 && r2 <= 0xDFFF) r = 0x10000 + ((r - 0xD800) << 10) + (r2 - 0xDC00);
                        else { i = save; r = 0xFFFD; }
                    } else if (r >= 0xD800 && r <= 0xDFFF) {
                        r = 0xFFFD;  // Go replaces lone surrogates
                    }
                    put_utf8(*o, r);
                    break;
                }
                default: return fail("bad escape");
            }
        }
        return fail("unterminated string");
    }
    ValuePtr value() {
        if (++depth > 512) { fail("nesting too deep"); return nullptr; }
        ws();
        ValuePtr v = std::make_shared<Value>();
        if (i >= t.size()) { fail("unexpected end"); return nullptr; }
        const char c = t[i];
        if (c == '{') {
            v->kind = Value::Object;
            ++i; ws();
            if (i < t.size() && t[i] == '}') { ++i; --depth; return v; }
            for (;;) {
                ws();
                std::string k;
                if (!string(&k)) return nullptr;
                ws();
                if (i >= t.size() || t[i] != ':') { fail("expected ':'"); return nullptr; }
                ++i;
                ValuePtr e = value();
                if (!e) return nullptr;
                v->obj.emplace_back(std::move(k), e);
                ws();
                if (i < t.size() && t[i] == ',') { ++i; continue; }
                if (i < t.size() && t[i] == '}') { ++i; break; }
                fail("expected ',' or '}'");
                return nullptr;
            }
        } else if (c == '[') {
            v->kind = Value::Array;
            ++i; ws();
            if (i < t.size() && t[i] == ']') { ++i; --depth; return v; }
            for (;;) {
                ValuePtr e = value();
                if (!e) return nullptr;
                v->arr.push_back(e);
                ws();
                if (i < t.size() && t[i] == ',') { ++i; continue; }
                if (i < t.size() && t[i] == ']') { ++i; break; }
                fail("expected ',' or ']'");
                return nullptr;
            }
        } else if (c == '"') {
            v->kind = Value::String;
            if (!string(&v->str)) return nullptr;
        } else if (t.compare(i, 4, "true") == 0) { v->kind = Value::Bool; v->b = true; i += 4; }
        else if (t.compare(i, 5, "false") == 0) { v->kind = Value::Bool; v->b = false; i += 5; }
        else if (t.compare(i, 4, "null") == 0) { v->kind = Value::Null; i += 4; }
        else if (c == '-' || (c >= '0' && c <= '9')) {
            const size_t s = i;
            if (t[i] == '-') ++i;
            if (i >= t.size()) { fail("bad number"); return nullptr; }
            if (t[i] == '0') ++i;
            else if (t[i] >= '1' && t[i] <= '9') { while (i < t.size() && isdigit((unsigned char)t[i])) ++i; }
            else { fail("bad number"); return nullptr; }
            bool is_int = true;
            if (i < t.size() && t[i] == '.') {
                is_int = false; ++i;
                if (i >= t.size() || !isdigit((unsigned char)t[i])) { fail("bad number"); return nullptr; }
                while (i < t.size() && isdigit((unsigned char)t[i])) ++i;
            }
            if (i < t.size() && (t[i] == 'e' || t[i] == 'E')) {
                is_int = false; ++i;
                if (i < t.size() && (t[i] == '+' || t[i] == '-')) ++i;
                if (i >= t.size() || !isdigit((unsigned char)t[i])) { fail("bad number"); return nullptr; }
                while (i < t.size() && isdigit((unsigned char)t[i])) ++i;
            }
            v->kind = Value::Number;
            const std::string lit = t.substr(s, i - s);
            v->num = strtod(lit.c_str(), nullptr);
            if (is_int) {
                errno = 0;
                v->inum = strtoll(lit.c_str(), nullptr, 10);
                v->is_int = (errno == 0);
            }
        } else {
            fail("unexpected character");
            return nullptr;
        }
        --depth;
        return v;
    }
};
}  // namespace

ValuePtr parse(const std::string& text, std::string* err) {
    Parser p(text);
    ValuePtr v = p.value();
    if (v) {
        p.ws();
        if (p.i != text.size()) {
            p.fail("trailing data");
            v.reset();
        }
    }
    if (!v && err) *err = p.err;
    return v;
}

}  // namespace gojson
}  // namespace cro
