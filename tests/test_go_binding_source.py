"""The cgo binding is source only (no Go toolchain in the image), so nothing compiles it against include/croprobe.h.
This check keeps the two from drifting: every C function, constant and type the Go files name exists in the header,
every struct field they read (`r.<field>`, `o.<field>` on C structs) is a member of that struct, and each call passes as
many arguments as the prototype declares."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "croprobe.h")).read()
GO_FILES = sorted(glob.glob(os.path.join(ROOT, "composable-resource-operator_b200", "go", "internal", "cuda", "*.go")))
CGO_BUILTINS = {"int", "char", "size_t", "uint64_t", "uint32_t", "int32_t", "uint8_t", "free", "malloc", "GoString", "GoStringN", "CString",
                "GoBytes", "uint", "long", "ulong", "longlong", "ulonglong", "double"}


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def _struct_fields(name):
    m = re.search(r"typedef struct\s+\w*\s*\{(.*?)\}\s*%s\s*;" % re.escape(name), _strip_comments(HEADER), re.S)
    assert m, name
    return set(re.findall(r"(\w+)\s*(?:\[[^\]]*\])*\s*;", m.group(1)))


def _prototype_arity(fn):
    m = re.search(r"\b%s\s*\(([^;{]*)\)\s*;" % re.escape(fn), _strip_comments(HEADER), re.S)
    assert m, "not declared in croprobe.h: " + fn
    args = m.group(1).strip()
    return 0 if args in ("", "void") else args.count(",") + 1


def _call_arity(src, pos):
    """number of top-level arguments of the call whose '(' is at src[pos]"""
    depth, n, i, seen = 0, 0, pos, False
    while True:
        c = src[i]
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
            if depth == 0:
                return n + (1 if seen else 0)
        elif c == "," and depth == 1:
            n += 1
        elif depth >= 1 and not c.isspace():
            seen = True
        i += 1


def test_go_files_exist():
    assert [os.path.basename(f) for f in GO_FILES] == ["probe.go", "token.go"]


def test_every_c_identifier_the_go_source_names_is_in_the_header():
    for path in GO_FILES:
        src = _strip_comments(open(path).read().split('import "C"', 1)[1])
        for ident in sorted(set(re.findall(r"\bC\.(\w+)", src))):
            if ident in CGO_BUILTINS:
                continue
            assert re.search(r"\b%s\b" % re.escape(ident), HEADER), (os.path.basename(path), ident)
        for m in re.finditer(r"\bC\.(cro_\w+)\(", src):
            fn = m.group(1)
            assert _call_arity(src, m.end() - 1) == _prototype_arity(fn), (os.path.basename(path), fn)


def test_struct_fields_read_by_the_go_source_exist():
    src = _strip_comments(open(GO_FILES[0]).read())
    body = src[src.index("func convert("):]
    body = body[:body.index("\n}\n")]
    result_fields = _struct_fields("cro_probe_result")
    used = set(re.findall(r"\br\.(\w+)", body))
    assert used and used <= result_fields, used - result_fields
    # options written in NewContext
    new_ctx = src[src.index("func NewContext("):]
    new_ctx = new_ctx[:new_ctx.index("\n}\n")]
    opts_fields = _struct_fields("cro_opts")
    written = set(re.findall(r"\bopts\.(\w+)\s*=", new_ctx)) | set(re.findall(r"\bo\.(\w+)\s*=", new_ctx))
    assert written <= opts_fields | set(), written - opts_fields


def test_preamble_includes_the_header_and_links_the_library():
    for path in GO_FILES:
        pre = open(path).read().split('import "C"', 1)[0]
        assert '#include "croprobe.h"' in pre and "-lcroprobe" in pre, path
