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
namespace {
// encoding/json matches an input key to a struct field exactly or under Unicode simple case folding
// (bytes.EqualFold).  Every field name we look up is ASCII, and the only non-ASCII runes that fold onto
// ASCII letters are U+017F (long s) and U+212A (Kelvin sign).
std::string foldKey(const std::string& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        const unsigned char c = (unsigned char)s[i];
        if (c >= 'A' && c <= 'Z') o.push_back((char)(c + 32));
        else if (c == 0xC5 && i + 1 < s.size() && (unsigned char)s[i + 1] == 0xBF) { o.push_back('s'); ++i; }
        else if (c == 0xE2 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x84 && (unsigned char)s[i + 2] == 0xAA) { o.push_back('k'); i += 2; }
        else o.push_back((char)c);
    }
    return o;
}
}  // namespace

// The member a Go struct field tagged `k` would receive: keys are taken in input order, each one that names
// the field (exactly or case-folded) overwrites the previous, so the LAST such key wins.
const Value* Value::get(const std::string& k) const {
    const Value* found = nullptr;
    std::string folded;
    for (const auto& kv : obj) {
        if (kv.first == k) { found = kv.second.get(); continue; }
        if (kv.first.size() < k.size()) continue;             // folding never lengthens: long s / Kelvin shrink
        if (folded.empty()) folded = foldKey(k);
        if (foldKey(kv.first) == folded) found = kv.second.get();
    }
    return found;
}
int MatchField(const std::string& key, const char* const* tags, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (key == tags[i]) return (int)i;
    const std::string folded = foldKey(key);
    for (size_t i = 0; i < n; ++i)
        if (foldKey(tags[i]) == folded) return (int)i;
    return -1;
}

std::string MismatchKind(const Value& v, const std::string& text, bool intTarget) {
    switch (v.kind) {
        case Value::Object: return "object";
        case Value::Array: return "array";
        case Value::String: return "string";
        case Value::Bool: return "bool";
        case Value::Number: return intTarget ? "number " + text.substr(v.raw_begin, v.raw_end - v.raw_begin) : "number";
        default: return "null";
    }
}

std::string DecodeFlat(const Value& root, const std::string& text, const char* structName, const FlatField* fields,
                       size_t n, std::map<std::string, std::string>* strs, std::map<std::string, long long>* ints) {
    std::string first;
    if (root.kind != Value::Object) return first;
    std::vector<const char*> tags;
    for (size_t i = 0; i < n; ++i) tags.push_back(fields[i].tag);
    for (const auto& kv : root.obj) {
        const int fi = MatchField(kv.first, tags.data(), n);
        if (fi < 0) continue;
        const FlatField* f = &fields[fi];
        const Value& v = *kv.second;
        if (v.kind == Value::Null) continue;
        if (f->type == 's' && v.kind == Value::String) { (*strs)[f->tag] = v.str; continue; }
        if (f->type == 'i' && v.kind == Value::Number && v.is_int) { (*ints)[f->tag] = v.inum; continue; }
        if (!first.empty()) continue;
        first = "json: cannot unmarshal " + MismatchKind(v, text, f->type == 'i') + " into Go struct field " + structName + "." +
                f->tag + " of type " + (f->type == 's' ? "string" : "int64");
    }
    return first;
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
constexpr int kMaxTreeDepth = 512;       // recursion bound of the tree builder (and of ~Value)
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
            if (c < 0x80) {
                if (c != '\\') { o->push_back((char)c); ++i; continue; }
            } else {
                // unquote() runs utf8.DecodeRune: a well-formed sequence is kept, anything else — stray continuation
                // byte, overlong form, encoded surrogate, > U+10FFFF, truncated tail — costs ONE byte and becomes U+FFFD
                unsigned rune = 0;
                const size_t n = utf8_len(t, i, &rune);
                if (n) { o->append(t, i, n); i += n; }
                else { put_utf8(*o, 0xFFFD); ++i; }
                continue;
            }
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
    // Steps over one container of already validated text without recursing (strings may hold brackets).
    void skip_container() {
        size_t open = 0;
        for (; i < t.size(); ++i) {
            const char c = t[i];
            if (c == '"') {
                for (++i; i < t.size() && t[i] != '"'; ++i)
                    if (t[i] == '\\') ++i;
            } else if (c == '{' || c == '[') {
                ++open;
            } else if (c == '}' || c == ']') {
                if (--open == 0) { ++i; return; }
            }
        }
    }
    ValuePtr value() {
        ++depth;
        ws();
        ValuePtr v = std::make_shared<Value>();
        v->raw_begin = i;
        if (i >= t.size()) { fail("unexpected end"); return nullptr; }
        const char c = t[i];
        if (depth > kMaxTreeDepth && (c == '{' || c == '[')) {
            // Deeper than any wire struct reaches (they end near depth 12): whatever sits here can only land in an
            // unknown field, a json.RawMessage or a map[string]any — none of which looks inside.  Kept as an empty
            // container that remembers its text span; parse() has already run Go's scanner over the whole input, so
            // the text is valid and Go's own limit (10000, "exceeded max depth") has been enforced there.
            v->kind = c == '{' ? Value::Object : Value::Array;
            skip_container();
            --depth;
            v->raw_end = i;
            return v;
        }
        if (c == '{') {
            v->kind = Value::Object;
            ++i; ws();
            if (i < t.size() && t[i] == '}') { ++i; --depth; v->raw_end = i; return v; }
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
            if (i < t.size() && t[i] == ']') { ++i; --depth; v->raw_end = i; return v; }
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
        v->raw_end = i;
        return v;
    }
};
}  // namespace

// ---- Go's validity scanner ----------------------------------------------------
// json.Unmarshal runs checkValid over the whole input before it decodes anything,
// so the text of every "invalid character ..." error the reference surfaces comes
// from this byte-at-a-time state machine (Go stdlib encoding/json/scanner.go,
// go1.24 per the reference's go.mod; restated from its published behaviour).
namespace {
enum class St {
    BeginValue, BeginValueOrEmpty, BeginStringOrEmpty, BeginString, EndValue, EndTop,
    InString, InStringEsc, EscU, EscU1, EscU12, EscU123,
    Neg, Num1, Num0, Dot, Dot0, E, ESign, E0,
    T, Tr, Tru, F, Fa, Fal, Fals, N, Nu, Nul
};
enum class Ps : unsigned char { ObjectKey, ObjectValue, ArrayValue };

std::string quoteChar(unsigned char c) {
    if (c == '\'') return "'\\''";
    if (c == '"') return "'\"'";
    // strconv.Quote(string(rune(c))) without its double quotes
    std::string s = "'";
    char buf[8];
    switch (c) {
        case '\a': s += "\\a"; break;
        case '\b': s += "\\b"; break;
        case '\f': s += "\\f"; break;
        case '\n': s += "\\n"; break;
        case '\r': s += "\\r"; break;
        case '\t': s += "\\t"; break;
        case '\v': s += "\\v"; break;
        case '\\': s += "\\\\"; break;
        default:
            if (c < 0x20 || c == 0x7f) { snprintf(buf, sizeof buf, "\\x%02x", c); s += buf; }
            else if (c < 0x80) s.push_back((char)c);
            else if (c < 0xa1 || c == 0xad) { snprintf(buf, sizeof buf, "\\u%04x", c); s += buf; }  // not IsPrint
            else { s.push_back((char)(0xC0 | (c >> 6))); s.push_back((char)(0x80 | (c & 0x3F))); }   // U+00A1..U+00FF
    }
    return s + "'";
}
inline bool isSpaceByte(unsigned char c) { return c <= ' ' && (c == ' ' || c == '\t' || c == '\r' || c == '\n'); }
inline bool isHex(unsigned char c) {
    return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F');
}
}  // namespace

std::string SyntaxError(const std::string& text) {
    constexpr size_t kMaxNestingDepth = 10000;
    St st = St::BeginValue;
    std::vector<Ps> stack;
    bool endTop = false;
    std::string err;
    auto bad = [&](unsigned char c, const char* context) {
        err = "invalid character " + quoteChar(c) + " " + context;
    };
    auto pop = [&]() {
        stack.pop_back();
        if (stack.empty()) { st = St::EndTop; endTop = true; }
        else st = St::EndValue;
    };
    // one step of the machine; `again` re-dispatches the same byte in a new state
    auto step = [&](unsigned char c) {
        for (;;) {
            switch (st) {
                case St::BeginValueOrEmpty:
                    if (isSpaceByte(c)) return;
                    if (c == ']') { st = St::EndValue; continue; }
                    st = St::BeginValue;
                    continue;
                case St::BeginValue:
                    if (isSpaceByte(c)) return;
                    switch (c) {
                        case '{':
                            st = St::BeginStringOrEmpty;
                            stack.push_back(Ps::ObjectKey);
                            if (stack.size() > kMaxNestingDepth) bad(c, "exceeded max depth");
                            return;
                        case '[':
                            st = St::BeginValueOrEmpty;
                            stack.push_back(Ps::ArrayValue);
                            if (stack.size() > kMaxNestingDepth) bad(c, "exceeded max depth");
                            return;
                        case '"': st = St::InString; return;
                        case '-': st = St::Neg; return;
                        case '0': st = St::Num0; return;
                        case 't': st = St::T; return;
                        case 'f': st = St::F; return;
                        case 'n': st = St::N; return;
                    }
                    if (c >= '1' && c <= '9') { st = St::Num1; return; }
                    bad(c, "looking for beginning of value");
                    return;
                case St::BeginStringOrEmpty:
                    if (isSpaceByte(c)) return;
                    if (c == '}') { stack.back() = Ps::ObjectValue; st = St::EndValue; continue; }
                    st = St::BeginString;
                    continue;
                case St::BeginString:
                    if (isSpaceByte(c)) return;
                    if (c == '"') { st = St::InString; return; }
                    bad(c, "looking for beginning of object key string");
                    return;
                case St::EndValue:
                    if (stack.empty()) { st = St::EndTop; endTop = true; continue; }
                    if (isSpaceByte(c)) { st = St::EndValue; return; }
                    switch (stack.back()) {
                        case Ps::ObjectKey:
                            if (c == ':') { stack.back() = Ps::ObjectValue; st = St::BeginValue; return; }
                            bad(c, "after object key");
                            return;
                        case Ps::ObjectValue:
                            if (c == ',') { stack.back() = Ps::ObjectKey; st = St::BeginString; return; }
                            if (c == '}') { pop(); return; }
                            bad(c, "after object key:value pair");
                            return;
                        case Ps::ArrayValue:
                            if (c == ',') { st = St::BeginValue; return; }
                            if (c == ']') { pop(); return; }
                            bad(c, "after array element");
                            return;
                    }
                    return;
                case St::EndTop:
                    if (!isSpaceByte(c)) bad(c, "after top-level value");
                    return;
                case St::InString:
                    if (c == '"') { st = St::EndValue; return; }
                    if (c == '\\') { st = St::InStringEsc; return; }
                    if (c < 0x20) bad(c, "in string literal");
                    return;
                case St::InStringEsc:
                    switch (c) {
                        case 'b': case 'f': case 'n': case 'r': case 't': case '\\': case '/': case '"':
                            st = St::InString; return;
                        case 'u': st = St::EscU; return;
                    }
                    bad(c, "in string escape code");
                    return;
                case St::EscU: case St::EscU1: case St::EscU12: case St::EscU123:
                    if (isHex(c)) {
                        st = st == St::EscU ? St::EscU1 : st == St::EscU1 ? St::EscU12 : st == St::EscU12 ? St::EscU123 : St::InString;
                        return;
                    }
                    bad(c, "in \\u hexadecimal character escape");
                    return;
                case St::Neg:
                    if (c == '0') { st = St::Num0; return; }
                    if (c >= '1' && c <= '9') { st = St::Num1; return; }
                    bad(c, "in numeric literal");
                    return;
                case St::Num1:
                    if (c >= '0' && c <= '9') return;
                    st = St::Num0;
                    continue;
                case St::Num0:
                    if (c == '.') { st = St::Dot; return; }
                    if (c == 'e' || c == 'E') { st = St::E; return; }
                    st = St::EndValue;
                    continue;
                case St::Dot:
                    if (c >= '0' && c <= '9') { st = St::Dot0; return; }
                    bad(c, "after decimal point in numeric literal");
                    return;
                case St::Dot0:
                    if (c >= '0' && c <= '9') return;
                    if (c == 'e' || c == 'E') { st = St::E; return; }
                    st = St::EndValue;
                    continue;
                case St::E:
                    if (c == '+' || c == '-') { st = St::ESign; return; }
                    st = St::ESign;
                    continue;
                case St::ESign:
                    if (c >= '0' && c <= '9') { st = St::E0; return; }
                    bad(c, "in exponent of numeric literal");
                    return;
                case St::E0:
                    if (c >= '0' && c <= '9') return;
                    st = St::EndValue;
                    continue;
#define CRO_LIT(state, want, next, text)                                             \
                case St::state:                                                      \
                    if (c == want) { st = St::next; return; }                        \
                    bad(c, "in literal " text " (expecting " #want ")");           \
                    return;
                CRO_LIT(T, 'r', Tr, "true") CRO_LIT(Tr, 'u', Tru, "true") CRO_LIT(Tru, 'e', EndValue, "true")
                CRO_LIT(F, 'a', Fa, "false") CRO_LIT(Fa, 'l', Fal, "false") CRO_LIT(Fal, 's', Fals, "false")
                CRO_LIT(Fals, 'e', EndValue, "false")
                CRO_LIT(N, 'u', Nu, "null") CRO_LIT(Nu, 'l', Nul, "null") CRO_LIT(Nul, 'l', EndValue, "null")
#undef CRO_LIT
            }
            return;
        }
    };
    for (unsigned char c : text) {
        step(c);
        if (!err.empty()) return err;
    }
    // eof(): one virtual space lets a pending number / literal finish
    if (endTop) return "";
    step(' ');
    if (!err.empty()) return err;
    if (endTop) return "";
    return "unexpected end of JSON input";
}

ValuePtr parse(const std::string& text, std::string* err) {
    const std::string syntax = SyntaxError(text);
    if (!syntax.empty()) {
        if (err) *err = syntax;
        return nullptr;
    }
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

bool rootOk(const ValuePtr& root, std::string* perr, const char* goType) {
    if (!root) return false;
    if (root->kind == Value::Object || root->kind == Value::Null) return true;
    static const char* kKind[] = {"null", "bool", "number", "string", "array", "object"};
    *perr = std::string("json: cannot unmarshal ") + kKind[root->kind] + " into Go value of type " + goType;
    return false;
}

namespace {
struct TypeWalk {
    const std::string& text;
    std::string first;
    std::vector<std::string> stack;                 // errorContext.FieldStack
    const GoType* strct = nullptr;                  // errorContext.Struct

    void save(const Value& v, const GoType& t) {
        if (!first.empty()) return;
        const std::string what = MismatchKind(v, text, t.kind == GoType::Int);
        if (!strct && stack.empty()) {
            first = "json: cannot unmarshal " + what + " into Go value of type " + t.name;
            return;
        }
        std::string path;
        for (size_t i = 0; i < stack.size(); ++i) path += (i ? "." : "") + stack[i];
        first = "json: cannot unmarshal " + what + " into Go struct field " + (strct ? strct->structName : std::string()) + "." + path +
                " of type " + t.name;
    }

    void walk(const Value& v, const GoType& t) {
        if (v.kind == Value::Null || t.kind == GoType::RawMessage) return;
        switch (t.kind) {
            case GoType::String: if (v.kind != Value::String) save(v, t); return;
            case GoType::Bool: if (v.kind != Value::Bool) save(v, t); return;
            case GoType::Int: if (v.kind != Value::Number || !v.is_int) save(v, t); return;
            case GoType::MapOfAny: if (v.kind != Value::Object) save(v, t); return;
            case GoType::Slice:
                if (v.kind != Value::Array) { save(v, t); return; }
                for (const auto& e : v.arr) walk(*e, *t.elem);
                return;
            case GoType::Struct: {
                if (v.kind != Value::Object) { save(v, t); return; }
                std::vector<const char*> tags;
                for (const auto& f : t.fields) tags.push_back(f.first.c_str());
                for (const auto& kv : v.obj) {
                    const int fi = MatchField(kv.first, tags.data(), tags.size());
                    if (fi < 0) continue;
                    const GoType* outer = strct;
                    stack.push_back(t.fields[fi].first);
                    strct = &t;
                    walk(*kv.second, *t.fields[fi].second);
                    stack.pop_back();
                    strct = outer;
                }
                return;
            }
            default: return;
        }
    }
};
}  // namespace

std::string TypeMismatch(const Value& root, const std::string& text, const GoType& t) {
    TypeWalk w{text};
    w.walk(root, t);
    return w.first;
}

namespace {
ValuePtr zeroOf(const GoType& t) {
    ValuePtr v = std::make_shared<Value>();
    switch (t.kind) {
        case GoType::String: v->kind = Value::String; break;
        case GoType::Int: v->kind = Value::Number; v->is_int = true; break;
        case GoType::Bool: v->kind = Value::Bool; break;
        case GoType::Struct: v->kind = Value::Object; break;
        default: v->kind = Value::Null; break;          // nil slice / map / RawMessage
    }
    return v;
}

// dst: what the Go variable holds so far (nullptr = zero value).  Nodes reachable from dst are either leaves shared
// with the parse tree (never written) or containers made here.
void mergeInto(ValuePtr& dst, const ValuePtr& src, const GoType& t) {
    if (src->kind == Value::Null) return;
    switch (t.kind) {
        case GoType::String: if (src->kind == Value::String) dst = src; return;
        case GoType::Bool: if (src->kind == Value::Bool) dst = src; return;
        case GoType::Int: if (src->kind == Value::Number && src->is_int) dst = src; return;
        case GoType::RawMessage: dst = src; return;
        case GoType::MapOfAny: {
            if (src->kind != Value::Object) return;
            ValuePtr m = std::make_shared<Value>();
            m->kind = Value::Object;
            if (dst && dst->kind == Value::Object) m->obj = dst->obj;           // an existing map is kept and written into
            for (const auto& kv : src->obj) {
                bool hit = false;
                for (auto& e : m->obj)
                    if (e.first == kv.first) { e.second = kv.second; hit = true; }
                if (!hit) m->obj.push_back(kv);
            }
            dst = m;
            return;
        }
        case GoType::Slice: {
            if (src->kind != Value::Array) return;
            ValuePtr a = std::make_shared<Value>();
            a->kind = Value::Array;
            for (size_t k = 0; k < src->arr.size(); ++k) {
                ValuePtr e = (dst && dst->kind == Value::Array && k < dst->arr.size()) ? dst->arr[k] : nullptr;
                if (e && e->kind == Value::Object) {            // a container made by an earlier pass: copy before writing
                    ValuePtr c = std::make_shared<Value>(*e);
                    e = c;
                }
                mergeInto(e, src->arr[k], *t.elem);
                a->arr.push_back(e ? e : zeroOf(*t.elem));
            }
            dst = a;
            return;
        }
        case GoType::Struct: {
            if (src->kind != Value::Object) return;
            ValuePtr o = std::make_shared<Value>();
            o->kind = Value::Object;
            if (dst && dst->kind == Value::Object) o->obj = dst->obj;
            std::vector<const char*> tags;
            for (const auto& f : t.fields) tags.push_back(f.first.c_str());
            for (const auto& kv : src->obj) {
                const int fi = MatchField(kv.first, tags.data(), tags.size());
                if (fi < 0) continue;
                const std::string& tag = t.fields[fi].first;
                ValuePtr* slot = nullptr;
                for (auto& e : o->obj)
                    if (e.first == tag) slot = &e.second;
                ValuePtr cur = slot ? *slot : nullptr;
                mergeInto(cur, kv.second, *t.fields[fi].second);
                if (!cur) continue;                              // nothing landed (null, or a skipped mismatch)
                if (slot) *slot = cur;
                else o->obj.emplace_back(tag, cur);
            }
            dst = o;
            return;
        }
    }
}
}  // namespace

ValuePtr DecodeAs(const ValuePtr& root, const std::string& text, const GoType& t, std::string* perr) {
    if (!decodesInto(root, text, t, perr)) return nullptr;
    ValuePtr out;
    mergeInto(out, root, t);
    return out ? out : zeroOf(t);
}

bool decodesInto(const ValuePtr& root, const std::string& text, const GoType& t, std::string* perr) {
    if (!root) return false;                        // *perr already holds the syntax error
    const std::string e = TypeMismatch(*root, text, t);
    if (e.empty()) return true;
    *perr = e;
    return false;
}

}  // namespace gojson
}  // namespace cro
