// gojson.hpp — byte-compatible writer for Go's encoding/json (go1.24, HTML
// escaping on, as json.Marshal uses it) plus a small strict JSON reader.
//
// The reference marshals its wire structs with json.Marshal
// (internal/cdi/fti/fm/client.go:144,271; fti/cm/client.go:139,218;
// sunfish/client.go:78) and its CR status through the apimachinery JSON codec,
// which is encoding/json as well.  "Bit-exact CDI JSON" therefore means:
// struct-declaration field order, omitempty rules, sorted map keys, and the
// stdlib's string escaping table — all restated here.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace cro {
namespace gojson {

// Appends the Go-escaped, quoted form of `s` to `out`.
void append_string(std::string& out, const std::string& s);

// Streaming writer for struct-shaped output.  Field order is the caller's
// (Go emits struct fields in declaration order).
class Writer {
public:
    Writer& begin_object();
    Writer& end_object();
    Writer& begin_array();
    Writer& end_array();
    Writer& key(const char* k);
    Writer& value(const std::string& s);
    Writer& value(const char* s);
    Writer& value(long long v);
    Writer& value(int v) { return value((long long)v); }
    Writer& value_u64(unsigned long long v);
    Writer& value(bool v);
    Writer& null();
    // key + value, skipped when the value is Go's zero value (omitempty).
    Writer& field_omitempty(const char* k, const std::string& s);
    Writer& field_omitempty(const char* k, bool v);
    Writer& field_omitempty(const char* k, long long v);
    Writer& field(const char* k, const std::string& s) { return key(k).value(s); }
    Writer& field(const char* k, long long v) { return key(k).value(v); }
    Writer& field(const char* k, int v) { return key(k).value((long long)v); }
    Writer& field(const char* k, bool v) { return key(k).value(v); }
    // map[string]string: Go sorts keys bytewise.
    Writer& string_map(const std::map<std::string, std::string>& m);
    Writer& raw(const std::string& json);
    const std::string& str() const { return out_; }
    std::string take() { return std::move(out_); }

private:
    void comma();
    std::string out_;
    std::vector<bool> first_;  // per nesting level: no element written yet
    bool after_key_ = false;
};

// ---- reader ---------------------------------------------------------------
struct Value;
using ValuePtr = std::shared_ptr<Value>;
struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    long long inum = 0;
    bool is_int = false;
    std::string str;
    std::vector<ValuePtr> arr;
    std::vector<std::pair<std::string, ValuePtr>> obj;  // insertion order
    size_t raw_begin = 0, raw_end = 0;                   // [begin, end) of this value in the parsed text (json.RawMessage)

    const Value* get(const std::string& k) const;  // Go's struct-field match: exact or case-folded key, last one wins
    std::string get_string(const std::string& k, const std::string& dflt = "") const;
    bool get_bool(const std::string& k, bool dflt = false) const;
    long long get_int(const std::string& k, long long dflt = 0) const;
};

// Go's json.Valid / checkValid as a message: "" when `text` is one valid JSON
// value, else the text of the *json.SyntaxError Unmarshal would return
// ("invalid character '<' looking for beginning of value", "unexpected end of
// JSON input", ...).  Pinned by the reference's own expected strings
// (composableresource_controller_test.go:1479,1674,1804,...).
std::string SyntaxError(const std::string& text);

// Returns nullptr and fills *err on malformed input; *err is SyntaxError(text) — which also enforces Go's
// nesting limit of 10000.  Containers nested deeper than 512 are kept as childless nodes that remember their
// text span: no wire struct reaches that deep, so only unknown fields / RawMessage / map[string]any see them.
ValuePtr parse(const std::string& text, std::string* err);

// Can a parse() result be decoded into a Go struct of type `goType`?  False with
// *perr = the message json.Unmarshal returns: the syntax error parse() left there,
// or "json: cannot unmarshal array into Go value of type api.X" when the top-level
// value is not an object.  JSON null decodes into the zero struct (true).
bool rootOk(const ValuePtr& root, std::string* perr, const char* goType);

// json.Unmarshal of an object into a struct whose fields are all `string` or `int64` (fti/token.go:40-54).
// Members are taken in input order; each is matched to a field by its tag (exactly, else case-folded), null
// leaves the field alone, a value of the wrong JSON type is skipped and the FIRST such mismatch is returned as
// go1.24's UnmarshalTypeError text ("json: cannot unmarshal number 1.5 into Go struct field token.expires_in of
// type int64") — decoding goes on, so later members still land.  (Mismatch wording: unpinned by the reference.)
// A Go type as json.Unmarshal sees it, for the typed walk below.  Only what the reference's wire structs use.
struct GoType {
    enum Kind { String, Int, Bool, Struct, Slice, RawMessage, MapOfAny } kind;
    std::string name;                                               // reflect.Type.String(): "int", "api.Condition", "[]api.ConditionItem"
    std::string structName;                                         // Struct: reflect.Type.Name(), "Condition"
    std::vector<std::pair<std::string, const GoType*>> fields;      // Struct: json tag -> type
    const GoType* elem = nullptr;                                   // Slice
};
// The error json.Unmarshal(text, &T{}) returns for a syntactically valid `root`: "" or go1.24's UnmarshalTypeError
// text of the FIRST value (input order) whose JSON type does not fit the Go field it lands in — "json: cannot
// unmarshal string into Go struct field GetMachineItem.data.machines.fabric_id of type int", or "... into Go value
// of type api.X" at the top.  null fits everything; unknown keys are skipped; keys fold as in Value::get.
// (The wording is go1.24's and has no reference vector; THAT a mismatch is an error is what matters to callers.)
std::string TypeMismatch(const Value& root, const std::string& text, const GoType& t);
// rootOk + TypeMismatch: can `root` (a parse() result) be decoded into T?  False with *perr = Unmarshal's message.
bool decodesInto(const ValuePtr& root, const std::string& text, const GoType& t, std::string* perr);
// What a zero T holds after json.Unmarshal(text, &v), as a canonical tree: a struct is an Object whose keys are
// exactly the tags of the members that were set; null left a field alone; a repeated struct member merged into
// what an earlier one stored; a repeated slice member was decoded over the earlier elements and truncated to the
// new length (decode.go object() / array()); a null slice element is the zero struct.  nullptr with *perr =
// Unmarshal's message when `root` is null (syntax error already in *perr) or some value does not fit T.
ValuePtr DecodeAs(const ValuePtr& root, const std::string& text, const GoType& t, std::string* perr);

// encoding/json's field lookup for one input key: index of the tag that equals it, else of the one equal under
// case folding, else -1.
int MatchField(const std::string& key, const char* const* tags, size_t n);
// The "%s" of UnmarshalTypeError for value v: "object" | "array" | "string" | "bool" | "number" — and, only where an
// integer field refused a number, "number <literal>".
std::string MismatchKind(const Value& v, const std::string& text, bool intTarget);

struct FlatField {
    const char* tag;
    char type;                      // 's' string, 'i' int64
};
std::string DecodeFlat(const Value& root, const std::string& text, const char* structName, const FlatField* fields,
                       size_t n, std::map<std::string, std::string>* strs, std::map<std::string, long long>* ints);

}  // namespace gojson
}  // namespace cro
