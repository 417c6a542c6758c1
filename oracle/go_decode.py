"""go_decode.py — how encoding/json (go1.24) fills the reference's wire structs, restated for the parity tests.

TEST INFRASTRUCTURE (part of the oracle): only tests/, __graft_entry__.smoke() and bench.py's CPU baseline use it.

What is restated: a span-keeping parse of syntactically valid JSON (members in input order, number literals kept);
the struct-field lookup (exact tag, else case-folded); the wire structs of internal/cdi/fti/fm/api/*.go and
internal/cdi/fti/cm/api/machine.go as type descriptions (TYPES); json.Unmarshal's UnmarshalTypeError for the first
value whose JSON type does not fit (type_mismatch); and the VALUE a zero struct holds afterwards (decode_as): null
leaves a field alone, a repeated struct member merges into what an earlier one stored, a repeated slice member is
decoded over the earlier elements and truncated to the new length (decode.go: object(), array()).

Parity status: pinned only indirectly — every golden reply of tests/golden/reference_entries.json decodes to what the
reference's entries expect.  The UnmarshalTypeError WORDING and the merge rules have no reference vector (the reference's
tests never send a mistyped or repeated member): parity unpinned for those; this file and csrc/gojson.cpp check each other."""
from __future__ import annotations

import json
import re
from typing import Dict, List, Optional, Tuple


def _fold(s: str) -> str:
    """encoding/json foldName: ASCII letters case-fold, U+017F -> s, U+212A -> k."""
    return "".join("s" if ch == "\u017f" else "k" if ch == "\u212a" else (ch.lower() if "A" <= ch <= "Z" else ch) for ch in s)


class _Node:
    """One JSON value of a syntactically valid text: kind in {"null","bool","number","string","array","object"}, its
    [begin, end) span, and v = python value | raw literal (number) | [nodes] | [(key, node)] in input order."""
    __slots__ = ("kind", "v", "b", "e")

    def __init__(self, kind, v, b, e):
        self.kind, self.v, self.b, self.e = kind, v, b, e


_WS = " \t\r\n"
_STR_RE = re.compile(r'"(?:[^"\\]|\\.)*"', re.S)
_NUM_RE = re.compile(r"-?(?:0|[1-9][0-9]*)(?:\.[0-9]+)?(?:[eE][+-]?[0-9]+)?")


def _parse_spans(s: str, i: int = 0) -> Tuple[_Node, int]:
    while s[i] in _WS:
        i += 1
    c = s[i]
    if c == "{":
        b, members = i, []
        i += 1
        while True:
            while s[i] in _WS:
                i += 1
            if s[i] == "}":
                return _Node("object", members, b, i + 1), i + 1
            if s[i] == ",":
                i += 1
                continue
            m = _STR_RE.match(s, i)
            key = json.loads(m.group(0), strict=False)
            i = m.end()
            while s[i] in _WS:
                i += 1
            val, i = _parse_spans(s, i + 1)            # s[i] == ":"
            members.append((key, val))
    if c == "[":
        b, items = i, []
        i += 1
        while True:
            while s[i] in _WS:
                i += 1
            if s[i] == "]":
                return _Node("array", items, b, i + 1), i + 1
            if s[i] == ",":
                i += 1
                continue
            val, i = _parse_spans(s, i)
            items.append(val)
    if c == '"':
        m = _STR_RE.match(s, i)
        return _Node("string", json.loads(m.group(0), strict=False), i, m.end()), m.end()
    for lit, val in (("true", True), ("false", False), ("null", None)):
        if s.startswith(lit, i):
            return _Node("null" if val is None else "bool", val, i, i + len(lit)), i + len(lit)
    m = _NUM_RE.match(s, i)
    return _Node("number", m.group(0), i, m.end()), m.end()


def _spans(data) -> Tuple[_Node, str]:
    s = data.decode("utf-8", "replace") if isinstance(data, bytes) else data
    return _parse_spans(s)[0], s


def _match_field(key: str, tags) -> Optional[str]:
    """encoding/json's field lookup: the exact tag, else the one equal under case folding."""
    if key in tags:
        return key
    f = _fold(key)
    for t in tags:
        if _fold(t) == f:
            return t
    return None


# ---- the wire structs as type descriptions (fm/api/*.go, cm/api/machine.go) ----
def _struct(_go_name: str, **fields):
    return ("struct", _go_name, dict(fields))


def _slice(elem):
    return ("slice", elem)


def _type_name(t) -> str:
    return t if isinstance(t, str) else "api." + t[1] if t[0] == "struct" else "[]" + _type_name(t[1])


_S, _I, _B = "string", "int", "bool"
_FM_COND = _struct("Condition", condition=_slice(_struct("ConditionItem", column=_S, operator=_S, value=_S)))
_FM_RES = dict(res_uuid=_S, res_name=_S, res_type=_S, res_status=_I, res_op_status=_S, res_serial_num=_S, res_spec=_FM_COND)
_FM_MACH = dict(fabric_uuid=_S, fabric_id=_I, mach_uuid=_S, mach_id=_I, mach_name=_S, tenant_uuid=_S)
TYPES = {
    "api.ScaleUpResponse": _struct("ScaleUpResponse", data=_struct("ScaleUpResponseData", machines=_slice(
        _struct("ScaleUpResponseMachineItem", resources=_slice(_struct("ScaleUpResponseResourceItem", **_FM_RES)), **_FM_MACH)))),
    "api.GetMachineResponse": _struct("GetMachineResponse", data=_struct("GetMachineData", machines=_slice(
        _struct("GetMachineItem", mach_status=_I, mach_status_detail=_S, resources=_slice(_struct("GetMachineResource", **_FM_RES)), **_FM_MACH)))),
    "api.MachineData": None,
}
_CM_DEVSPEC = _struct("DeviceResourceSpec", resspec_uuid=_S, productname=_S, model=_S, vendor=_S, removable=_B)
_CM_DETAIL = _struct("DeviceDetail", fabric_uuid=_S, fabric_id=_I, res_uuid=_S, fabr_gid=_S, res_type=_S, res_name=_S, res_status=_S,
                     res_op_status=_S, tenant_uuid=_S, mach_uuid=_S, resspecs=_slice(_CM_DEVSPEC))
_CM_DEVICE = _struct("Device", device_id=_S, status=_S, status_reason=_S, detail=_CM_DETAIL)
_CM_SELECTOR = _struct("Selector", version=_S, expression=_struct("Expression", conditions=_slice(
    _struct("Condition", column=_S, operator=_S, value=_S))))
_CM_SPEC = _struct("ResourceSpec", spec_uuid=_S, type=_S, min_resspec_count=_I, max_resspec_count=_I, device_count=_I,
                   selector=_CM_SELECTOR, devices=_slice(_CM_DEVICE))
_CM_MACHINE = _struct("Machine", uuid=_S, name=_S, status=_S, status_reason=_S, resspecs=_slice(_CM_SPEC))
TYPES["api.MachineData"] = _struct("MachineData", data=_struct("Data", tenant_uuid=_S, cluster=_struct(
    "Cluster", cluster_uuid=_S, machine=_CM_MACHINE)))


def type_mismatch(data, typ) -> str:
    """"" or the UnmarshalTypeError json.Unmarshal(data, &T{}) returns for syntactically valid data: the first value, in
    input order, whose JSON type does not fit the Go field it lands in (go1.24 wording: Struct = innermost struct's name,
    Field = dotted path of tags from the root; slices add nothing to the path)."""
    top, _ = _spans(data)
    first: List[str] = []

    def save(n: _Node, t, stack, strct):
        if first:
            return
        what = ("number " + n.v) if (n.kind == "number" and t == "int") else n.kind
        if strct is None and not stack:
            first.append("json: cannot unmarshal %s into Go value of type %s" % (what, _type_name(t)))
        else:
            first.append("json: cannot unmarshal %s into Go struct field %s.%s of type %s" % (what, strct or "", ".".join(stack), _type_name(t)))

    def walk(n: _Node, t, stack, strct):
        if n.kind == "null":
            return
        if t == "string":
            ok = n.kind == "string"
        elif t == "bool":
            ok = n.kind == "bool"
        elif t == "int":
            ok = _as_int(n) is not None
        elif t[0] == "slice":
            ok = n.kind == "array"
            if ok:
                for e in n.v:
                    walk(e, t[1], stack, strct)
        else:
            ok = n.kind == "object"
            if ok:
                for k, m in n.v:
                    tag = _match_field(k, t[2])
                    if tag is not None:
                        walk(m, t[2][tag], stack + [tag], t[1])
        if not ok:
            save(n, t, stack, strct)

    walk(top, typ, [], None)
    return first[0] if first else ""


def _as_int(n: _Node) -> Optional[int]:
    if n.kind == "number" and re.fullmatch(r"-?[0-9]+", n.v) and -2**63 <= int(n.v) < 2**63:
        return int(n.v)
    return None


def _zero(t):
    return "" if t == "string" else 0 if t == "int" else False if t == "bool" else {} if t[0] == "struct" else None


def decode_as(data, typ):
    """The value a zero T holds after json.Unmarshal(data, &v), as plain python (struct -> {tag: value} with only the
    members that were set, slice -> list or None when never set).  Members of the wrong JSON type are skipped, as the
    decoder skips them after recording its error (type_mismatch reports that error)."""
    top, _ = _spans(data)

    def merge(cur, n: _Node, t):
        if n.kind == "null":
            return cur
        if t == "string":
            return n.v if n.kind == "string" else cur
        if t == "bool":
            return n.v if n.kind == "bool" else cur
        if t == "int":
            return _as_int(n) if _as_int(n) is not None else cur
        if t[0] == "slice":
            if n.kind != "array":
                return cur
            old = cur if isinstance(cur, list) else []
            out = []
            for k, e in enumerate(n.v):            # element k is decoded over what an earlier array left at k
                out.append(merge(old[k] if k < len(old) else _zero(t[1]), e, t[1]))
            return out
        if n.kind != "object":
            return cur
        cur = dict(cur) if isinstance(cur, dict) else {}
        for k, m in n.v:
            tag = _match_field(k, t[2])
            if tag is not None:
                v = merge(cur.get(tag, _zero(t[2][tag])), m, t[2][tag])
                if v is not None:
                    cur[tag] = v
        return cur

    out = merge(_zero(typ), top, typ)
    return out if out is not None else _zero(typ)
