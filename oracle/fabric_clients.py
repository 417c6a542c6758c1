"""CPU restatement (TEST INFRASTRUCTURE — only tests/ may import this) of the reference's two FTI
CdiProvider clients and the pieces of Go's standard library whose error TEXT they surface.

  go_json_syntax_error   encoding/json checkValid (go1.24 per go.mod:3-5): the message of the
                         *json.SyntaxError Unmarshal returns; pinned by the reference's expected strings
                         "invalid character '<' looking for beginning of value" (…_test.go:1804 …) and
                         "invalid character 'h' in literal true (expecting 'r')" (:1674)
  FMClient / CMClient    internal/cdi/fti/fm/client.go:100-511, internal/cdi/fti/cm/client.go:107-509
  parse_rfc3339          time.Parse(time.RFC3339, s) incl. *time.ParseError text (:2863)
  restart_daemonset      internal/utils/nodes.go:35-76

Independent of csrc/provider.cpp, csrc/gojson.cpp and csrc/nodes.cpp (regex/table driven here,
hand-rolled state machines there); tests fuzz one against the other on the same scripted fabric."""
from __future__ import annotations

import json
import re
from typing import Dict, List, Optional, Tuple

import oracle as _o

ERR_ATTACHING = "device is attaching to the cluster"
ERR_DETACHING = "device is detaching from the cluster"


# --------------------------------------------------------------------------
# encoding/json scanner
# --------------------------------------------------------------------------
def _quote_char(c: int) -> str:
    if c == 0x27:
        return "'\\''"
    if c == 0x22:
        return "'\"'"
    named = {7: "\\a", 8: "\\b", 12: "\\f", 10: "\\n", 13: "\\r", 9: "\\t", 11: "\\v", 0x5C: "\\\\"}
    if c in named:
        return "'" + named[c] + "'"
    if c < 0x20 or c == 0x7F:
        return "'\\x%02x'" % c
    if c < 0x80:
        return "'" + chr(c) + "'"
    if c < 0xA1 or c == 0xAD:
        return "'\\u%04x'" % c
    return "'" + chr(c) + "'"


_LITERALS = {"t": "true", "f": "false", "n": "null"}


def go_json_syntax_error(data: bytes) -> str:
    """"" if `data` is one valid JSON value, else Go's SyntaxError text.  Recursive-descent with the
    scanner's error wording; the C++ side is the byte-at-a-time machine."""
    n = len(data)
    SPACE = b" \t\r\n"

    class Bad(Exception):
        pass

    def bad(i: int, ctx: str):
        raise Bad("invalid character " + _quote_char(data[i]) + " " + ctx)

    def eof():
        raise Bad("unexpected end of JSON input")

    def ws(i: int) -> int:
        while i < n and data[i] in SPACE:
            i += 1
        return i

    def string(i: int) -> int:        # data[i] == '"'
        i += 1
        while True:
            if i >= n:
                eof()
            c = data[i]
            if c == 0x22:
                return i + 1
            if c == 0x5C:
                i += 1
                if i >= n:            # the scanner feeds one virtual space at EOF
                    raise Bad("invalid character ' ' in string escape code")
                e = data[i]
                if e in b'bfnrt\\/"':
                    i += 1
                    continue
                if e == 0x75:
                    for k in range(1, 5):
                        if i + k >= n:
                            raise Bad("invalid character ' ' in \\u hexadecimal character escape")
                        if data[i + k] not in b"0123456789abcdefABCDEF":
                            bad(i + k, "in \\u hexadecimal character escape")
                    i += 5
                    continue
                bad(i, "in string escape code")
            if c < 0x20:
                bad(i, "in string literal")
            i += 1

    def number(i: int) -> int:
        if data[i] == 0x2D:
            i += 1
            if i >= n:
                raise Bad("invalid character ' ' in numeric literal")
            if not (0x30 <= data[i] <= 0x39):
                bad(i, "in numeric literal")
        if data[i] == 0x30:
            i += 1
        else:
            while i < n and 0x30 <= data[i] <= 0x39:
                i += 1
        if i < n and data[i] == 0x2E:
            i += 1
            if i >= n:
                raise Bad("invalid character ' ' after decimal point in numeric literal")
            if not (0x30 <= data[i] <= 0x39):
                bad(i, "after decimal point in numeric literal")
            while i < n and 0x30 <= data[i] <= 0x39:
                i += 1
        if i < n and data[i] in b"eE":
            i += 1
            if i < n and data[i] in b"+-":
                i += 1
            if i >= n:
                raise Bad("invalid character ' ' in exponent of numeric literal")
            if not (0x30 <= data[i] <= 0x39):
                bad(i, "in exponent of numeric literal")
            while i < n and 0x30 <= data[i] <= 0x39:
                i += 1
        return i

    def value(i: int, depth: int) -> int:
        i = ws(i)
        if i >= n:
            eof()
        c = data[i]
        if c == 0x7B:                                     # {
            if depth + 1 > 10000:
                bad(i, "exceeded max depth")
            i = ws(i + 1)
            if i >= n:
                eof()
            if data[i] == 0x7D:
                return i + 1
            while True:
                i = ws(i)
                if i >= n:
                    eof()
                if data[i] != 0x22:
                    bad(i, "looking for beginning of object key string")
                i = ws(string(i))
                if i >= n:
                    eof()
                if data[i] != 0x3A:
                    bad(i, "after object key")
                i = ws(value(i + 1, depth + 1))
                if i >= n:
                    eof()
                if data[i] == 0x2C:
                    i += 1
                    continue
                if data[i] == 0x7D:
                    return i + 1
                bad(i, "after object key:value pair")
        if c == 0x5B:                                     # [
            if depth + 1 > 10000:
                bad(i, "exceeded max depth")
            i = ws(i + 1)
            if i >= n:
                eof()
            if data[i] == 0x5D:
                return i + 1
            while True:
                i = ws(value(i, depth + 1))
                if i >= n:
                    eof()
                if data[i] == 0x2C:
                    i += 1
                    continue
                if data[i] == 0x5D:
                    return i + 1
                bad(i, "after array element")
        if c == 0x22:
            return string(i)
        if c == 0x2D or 0x30 <= c <= 0x39:
            return number(i)
        ch = chr(c)
        if ch in _LITERALS:
            word = _LITERALS[ch]
            for k in range(1, len(word)):
                if i + k >= n:
                    # the scanner feeds one virtual space at EOF
                    raise Bad("invalid character ' ' in literal %s (expecting '%s')" % (word, word[k]))
                if data[i + k] != ord(word[k]):
                    bad(i + k, "in literal %s (expecting '%s')" % (word, word[k]))
            return i + len(word)
        bad(i, "looking for beginning of value")

    try:
        import sys
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 50000))
        try:
            i = ws(value(0, 0))
        finally:
            sys.setrecursionlimit(old)
        if i < n:
            bad(i, "after top-level value")
        return ""
    except Bad as e:
        return str(e)


def unmarshal(data, go_type: str):
    """(value, err): json.Unmarshal into a struct of type go_type (objects and null only)."""
    raw = data if isinstance(data, bytes) else data.encode("utf-8", "surrogatepass")
    err = go_json_syntax_error(raw)
    if err:
        return None, err
    if go_type in TYPES:
        err = type_mismatch(data, TYPES[go_type])
        if err:
            return None, err
    v = _o.go_loads(raw.decode("utf-8", "replace") if isinstance(data, bytes) else data)   # invalid UTF-8 -> U+FFFD, as Go
    if v is None:
        return {}, ""
    if not isinstance(v, dict):
        kind = "array" if isinstance(v, list) else "string" if isinstance(v, str) else "bool" if isinstance(v, bool) else "number"
        return None, "json: cannot unmarshal %s into Go value of type %s" % (kind, go_type)
    return v, ""


# --------------------------------------------------------------------------
# the id_manager's answer -> token expiry or error (fti/token.go:96-175)
# --------------------------------------------------------------------------
_B64URL = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_"


def decode_base64_raw_url(s: str) -> Tuple[Optional[bytes], str]:
    """base64.RawURLEncoding.DecodeString: CR/LF skipped, no padding, non-strict trailing bits."""
    raw = s.encode("utf-8", "surrogatepass")
    acc: List[int] = []
    out = bytearray()
    for i, c in enumerate(raw):
        if c in (10, 13):
            continue
        v = _B64URL.find(chr(c)) if c < 128 else -1
        if v < 0:
            return None, "illegal base64 data at input byte %d" % i
        acc.append(v)
        if len(acc) == 4:
            out += bytes([(acc[0] << 2 | acc[1] >> 4) & 255, (acc[1] << 4 | acc[2] >> 2) & 255, (acc[2] << 6 | acc[3]) & 255])
            acc = []
    if len(acc) == 1:
        return None, "illegal base64 data at input byte %d" % (len(raw) - 1)
    if len(acc) >= 2:
        out.append((acc[0] << 2 | acc[1] >> 4) & 255)
    if len(acc) == 3:
        out.append((acc[1] << 4 | acc[2] >> 2) & 255)
    return bytes(out), ""


from go_decode import (TYPES, _Node, _WS, _as_int, _match_field, _spans, decode_as, type_mismatch)  # noqa: E402,F401


_TOKEN_FIELDS = {"access_token": "string", "expires_in": "int64", "refresh_expires_in": "int64", "refresh_token": "string",
                 "token_type": "string", "id_token": "string", "not-before-policy": "int64", "session_state": "string",
                 "scope": "string"}                                    # fti/token.go:40-50


def decode_flat(data, struct: str, fields: Dict[str, str]) -> Tuple[Dict, str]:
    """json.Unmarshal of syntactically valid `data` (object or null) into a struct of string / int64 fields:
    (values, first UnmarshalTypeError text).  Members in input order; a mismatch is skipped, decoding goes on."""
    top, _ = _spans(data)
    vals: Dict = {}
    first = ""
    if top.kind != "object":
        return vals, first
    for k, n in top.v:
        tag = _match_field(k, fields)
        if tag is None or n.kind == "null":
            continue
        typ = fields[tag]
        if typ == "string" and n.kind == "string":
            vals[tag] = n.v
            continue
        if typ == "int64" and _as_int(n) is not None:
            vals[tag] = _as_int(n)
            continue
        if first:
            continue
        what = ("number " + n.v) if (n.kind == "number" and typ == "int64") else n.kind
        first = "json: cannot unmarshal %s into Go struct field %s.%s of type %s" % (what, struct, tag, typ)
    return vals, first


def token_from_reply(t: Dict) -> Tuple[int, str]:
    """(expiry unix, "") or (0, the error CachedToken.Token returns)."""
    if t.get("secret_error"):
        return 0, t["secret_error"]                                   # token.go:98-101
    if t.get("transport_error"):
        return 0, t["transport_error"]                                # token.go:127-130
    status, body = int(t.get("status", 200)), t.get("body", "")
    if status != 200:
        return 0, "http returned code: %d, response body: %s" % (status, body)          # :138-140
    tok, err = unmarshal(body, "fti.token")
    if err:
        return 0, "failed to read id_manager response body into Token: " + err           # :143-145
    vals, err = decode_flat(body, "token", _TOKEN_FIELDS)
    if err:
        return 0, "failed to read id_manager response body into Token: " + err
    access = vals.get("access_token", "")
    parts = access.split(".")
    if len(parts) != 3:
        return 0, "invalid access token: " + access                                      # :156-159
    payload, err = decode_base64_raw_url(parts[1])
    if err:
        return 0, "failed to decode id_manager payload: " + err                          # :161-164
    claims, err = unmarshal(payload, "fti.accessToken")
    if not err:
        claims, err = decode_flat(payload, "accessToken", {"exp": "int64"})
    if err:
        return 0, "failed to unmarshal id_manager json: " + err                          # :166-169
    return claims.get("exp", 0), ""


# --------------------------------------------------------------------------
# the scripted fabric (same JSON the C harness takes under "fabric")
# --------------------------------------------------------------------------
class Fabric:
    def __init__(self, spec: Optional[Dict]):
        self.spec = spec or {}
        self.requests: List[Dict[str, str]] = []
        self.status_updates: List[str] = []
        self.now = 1735689600                      # 2025-01-01T00:00:00Z unless the caller sets it
        self.token_fetches = 0
        self._expiry: Optional[int] = None

    def do(self, method: str, path: str, query: str, body: str) -> Tuple[int, str, str]:
        self.requests.append({"method": method, "path": path, "query": query, "body": body})
        if self.spec.get("transport_error"):
            return 0, "", self.spec["transport_error"]
        for r in self.spec.get("http") or []:
            if r.get("method") and r["method"] != method:
                continue
            if isinstance(r.get("path"), str):
                if r["path"] != path:
                    continue
            elif r.get("path_contains", "") not in path:
                continue
            return int(r.get("status", 200)), r.get("body", ""), ""
        return 0, "", '%s "https://fabric/%s": no route in the scripted fabric' % (method, path)

    def token(self) -> str:
        """CachedToken.GetToken (fti/token.go:72-94): "" or the error text."""
        t = self.spec.get("token")
        if not isinstance(t, dict):
            return self.spec.get("token_error", "")
        if self._expiry is not None and self._expiry - 30 > self.now:      # leeway, token.go:68,78
            return ""
        self.token_fetches += 1
        exp, err = token_from_reply(t)
        if err:
            return "unable to rotate token: " + err
        self._expiry = exp
        return ""

    def _get(self, coll: str, resource: str, key: str, name: str):
        v = ((self.spec.get("objects") or {}).get(coll) or {}).get(key)
        if not isinstance(v, dict):
            return None, '%s "%s" not found' % (resource, name)
        return v, ""

    def node(self, name):
        return self._get("nodes", "nodes", name, name)

    def machine(self, ns, name):
        return self._get("metal3machines", "metal3machines.infrastructure.cluster.x-k8s.io", ns + "/" + name, name)

    def bmh(self, ns, name):
        return self._get("baremetalhosts", "baremetalhosts.metal3.io", ns + "/" + name, name)

    def node_names(self) -> List[str]:
        return list(((self.spec.get("objects") or {}).get("nodes") or {}).keys())

    def device_ids(self) -> List[str]:
        return [d for d in ((self.spec.get("objects") or {}).get("composable_resource_device_ids") or []) if isinstance(d, str)]


def _ann(obj: Dict, key: str) -> str:
    a = obj.get("annotations")
    v = a.get(key) if isinstance(a, dict) else None
    return v if isinstance(v, str) else ""


def machine_id_from_annotations(f: Fabric, node_name: str) -> Tuple[str, str]:
    """cm/client.go:348-388 (== fm/client.go:415-450)."""
    node, err = f.node(node_name)
    if err:
        return "", err
    info = _ann(node, "machine.openshift.io/machine")
    parts = info.split("/")
    if len(parts) != 2:
        return "", "failed to get annotation 'machine.openshift.io/machine' from Node %s, now is '%s'" % (node_name, info)
    machine, err = f.machine(parts[0], parts[1])
    if err:
        return "", err
    binfo = _ann(machine, "metal3.io/BareMetalHost")
    bparts = binfo.split("/")
    if len(bparts) != 2:
        return "", "failed to get annotation 'metal3.io/BareMetalHost' from Machine %s, now is '%s'" % (parts[1], binfo)
    bmh, err = f.bmh(bparts[0], bparts[1])
    if err:
        return "", err
    mid = _ann(bmh, "cluster-manager.cdi.io/machine")
    if not isinstance(bmh.get("annotations"), dict) or mid == "":
        return "", "failed to get annotation 'cluster-manager.cdi.io/machine' from BareMetalHost %s, now is '%s'" % (bparts[1], mid)
    return mid, ""


def _error_body(data: str, fm: bool):
    """(status, code, message, err) of Unmarshal(data, &api.ErrorBody{})."""
    v, err = unmarshal(data, "api.ErrorBody")
    if err:
        return 0, "", "", err
    top, text = _spans(data)
    state = {"first": "", "status": 0, "code": "", "message": ""}

    def mismatch(n: _Node, where: str, typ: str, literal: bool):
        if not state["first"]:
            what = ("number " + n.v) if (n.kind == "number" and literal) else n.kind
            state["first"] = "json: cannot unmarshal %s into Go struct field %s of type %s" % (what, where, typ)

    def detail(d: _Node):
        for k, n in d.v:
            tag = _match_field(k, ("code", "message", "data") if fm else ("code", "message"))
            if tag is None:
                continue
            if tag == "message" and fm:                      # json.RawMessage: the member's own text, null included
                state["message"] = "" if n.kind == "null" else n.v if n.kind == "string" else text[n.b:n.e].strip(_WS)
            elif n.kind == "null":
                continue
            elif tag == "data":
                if n.kind != "object":
                    mismatch(n, "ErrorDetail.detail.data", "map[string]interface {}", False)
            elif n.kind == "string":
                state[tag] = n.v
            else:
                mismatch(n, "ErrorDetail.detail." + tag, "string", False)

    if top.kind == "object":
        for k, n in top.v:                                   # members in input order; the first mismatch is the one reported
            tag = _match_field(k, ("status", "detail"))
            if tag is None or n.kind == "null":
                continue
            if tag == "status":
                if _as_int(n) is not None:
                    state["status"] = _as_int(n)
                else:
                    mismatch(n, "ErrorBody.status", "int", True)
            elif n.kind == "object":
                detail(n)
            else:
                mismatch(n, "ErrorBody.detail", "api.ErrorDetail", False)
    return state["status"], state["code"], state["message"], state["first"]


def fm_error(what: str, body: str) -> str:
    status, code, message, err = _error_body(body, True)
    if err:
        subject = "scaledown" if what == "scaledown" else "FM " + what
        return "failed to unmarshal %s error response body into errBody. Original error: %s" % (subject, err)
    return "failed to process FM %s request. FM returned code: '%s', error message: '%s'" % (what, code, message)


def cm_error(what: str, body: str) -> str:
    status, code, message, err = _error_body(body, False)
    if err:
        return "failed to unmarshal CM %s error response body into errBody. Original error: %s" % (what, err)
    if what == "scaledown":
        return "failed to process CM scaledown request. http returned status: %d, cm return code: %s, error message: %s" % (status, code, message)
    return "failed to process CM %s request. http returned status: '%d', cm return code: '%s', error message: '%s'" % (what, status, code, message)


def _query_escape(s: str) -> str:      # net/url.QueryEscape
    from urllib.parse import quote_plus
    return quote_plus(s, safe="-_.~")


class _Client:
    def __init__(self, fabric: Fabric, tenant: str, cluster: str):
        self.f, self.tenant, self.cluster = fabric, tenant, cluster


class FMClient(_Client):
    def machine_id(self, node_name: str) -> Tuple[str, str]:
        if self.cluster != "":
            return machine_id_from_annotations(self.f, node_name)
        node, err = self.f.node(node_name)
        if err:
            return "", err
        pid = node.get("provider_id") if isinstance(node.get("provider_id"), str) else ""
        if not pid.startswith("fsas-cdi://"):
            return "", "invalid format: expected 'fsas-cdi://machineUUID', now is '%s'" % pid
        return pid[len("fsas-cdi://"):], ""

    def machine_info(self, mid: str) -> Tuple[str, str]:
        if self.f.token():
            return "", self.f.token()
        st, body, terr = self.f.do("GET", "fabric_manager/api/v1/machines/" + mid, "tenant_uuid=" + _query_escape(self.tenant), "")
        if terr:
            return "", terr
        if st != 200:
            return "", fm_error("get", body)
        _, err = unmarshal(body, "api.GetMachineResponse")
        if err:
            return "", "failed to unmarshal FM get machine response body into machineData: " + err
        return body, ""

    def add(self, name, typ, model, node) -> Tuple[str, str, str]:
        mid, err = self.machine_id(node)
        if err:
            return "", "", err
        if self.f.token():
            return "", "", self.f.token()
        st, body, terr = self.f.do("PATCH", "fabric_manager/api/v1/machines/%s/update" % mid, "tenant_uuid=" + _query_escape(self.tenant),
                                   _o.emit_fm_scale_up(self.tenant, mid, typ, model))
        if terr:
            return "", "", terr
        if st != 200:
            return "", "", fm_error("scaleup", body)
        _, err = unmarshal(body, "api.ScaleUpResponse")
        if err:
            return "", "", "failed to unmarshal FM scaleup response body into scaleUpResponse. Original error: " + err
        return _o.fm_scale_up_response_to_ids(body if _o.go_loads(body) is not None else "{}", name, typ, model)

    def remove(self, typ, node, cdi_device_id) -> str:
        mid, err = self.machine_id(node)
        if err:
            return err
        body, err = self.machine_info(mid)
        if err:
            return err
        machines = ((_o.go_struct_loads(body, "api.GetMachineResponse") or {}).get("data") or {}).get("machines") or []
        if not machines:
            return "runtime error: index out of range [0] with length 0"
        if not any(r.get("res_type", "") == typ and r.get("res_uuid", "") == cdi_device_id for r in machines[0].get("resources") or []):
            return ""
        if self.f.token():
            return self.f.token()
        st, rbody, terr = self.f.do("DELETE", "fabric_manager/api/v1/machines/%s/update" % mid, "tenant_uuid=" + _query_escape(self.tenant),
                                    _o.emit_fm_scale_down(self.tenant, mid, typ, cdi_device_id))
        if terr:
            return terr
        if st not in (200, 204):
            return fm_error("scaledown", rbody)
        return ""

    def check(self, typ, model, node, device_id) -> str:
        mid, err = self.machine_id(node)
        if err:
            return err
        body, err = self.machine_info(mid)
        if err:
            return err
        return _o.fabric_check_resource("fm", body if _o.go_loads(body) is not None else "{}", typ, model, device_id)

    def resources(self) -> Tuple[List[Dict[str, str]], str]:
        out: List[Dict[str, str]] = []
        for n in self.f.node_names():
            mid, err = self.machine_id(n)
            if err:
                continue
            body, err = self.machine_info(mid)
            if err:
                continue
            out += _o.fabric_get_resources("fm", body if _o.go_loads(body) is not None else "{}", n, mid)
        return out, ""


class CMClient(_Client):
    def _path(self, mid: str) -> str:
        return "cluster_manager/cluster_autoscaler/v3/tenants/%s/clusters/%s/machines/%s" % (self.tenant, self.cluster, mid)

    def machine_info(self, mid: str) -> Tuple[str, str]:
        if self.f.token():
            return "", self.f.token()
        st, body, terr = self.f.do("GET", self._path(mid), "", "")
        if terr:
            return "", terr
        if st != 200:
            return "", cm_error("get", body)
        _, err = unmarshal(body, "api.MachineData")
        if err:
            return "", "failed to unmarshal CM get machine response body into machineData: " + err
        return body if _o.go_loads(body) is not None else "{}", ""

    def add(self, name, typ, model, node) -> Tuple[str, str, str]:
        mid, err = machine_id_from_annotations(self.f, node)
        if err:
            return "", "", err
        body, err = self.machine_info(mid)
        if err:
            return "", "", err
        spec_uuid, count, dev, res, cerr = _o.cm_check_adding_resources(body, self.f.device_ids(), typ, model)
        if dev != "":
            return dev, res, cerr
        if self.f.token():
            return "", "", self.f.token()
        st, rbody, terr = self.f.do("POST", self._path(mid) + "/actions/resize", "", _o.emit_cm_scale_up(spec_uuid, count + 1))
        if terr:
            return "", "", terr
        if st != 200:
            return "", "", cm_error("scaleup", rbody)
        return "", "", ERR_ATTACHING

    def remove(self, typ, model, node, device_id) -> Tuple[str, Optional[str]]:
        """(err, Status.Error recorded by the REMOVE_FAILED branch or None)."""
        mid, err = machine_id_from_annotations(self.f, node)
        if err:
            return err, None
        body, err = self.machine_info(mid)
        if err:
            return err, None
        spec_uuid, count, reason = "", 0, None
        for s in ((((_o.go_struct_loads(body, "api.MachineData").get("data") or {}).get("cluster") or {}).get("machine") or {}).get("resspecs")) or []:
            if s.get("type", "") != typ:
                continue
            conds = (((s.get("selector") or {}).get("expression") or {}).get("conditions")) or []
            if not any(c.get("column") == "model" and c.get("operator") == "eq" and c.get("value") == model for c in conds):
                continue
            for d in s.get("devices") or []:
                if d.get("device_id", "") == device_id:
                    spec_uuid, count = s.get("spec_uuid", ""), int(s.get("device_count", 0) or 0)
                    if d.get("status", "") == "REMOVE_FAILED":
                        reason = d.get("status_reason", "")
                    break
            break
        recorded = None
        if reason is not None:
            recorded = reason
            upd = (self.f.spec.get("objects") or {}).get("status_update_error", "")
            if upd:
                return upd, recorded
        if spec_uuid == "":
            return "", recorded
        if self.f.token():
            return self.f.token(), recorded
        st, rbody, terr = self.f.do("POST", self._path(mid) + "/actions/resize", "", _o.emit_cm_scale_down(spec_uuid, count - 1, device_id))
        if terr:
            return terr, recorded
        if st != 200:
            return cm_error("scaledown", rbody), recorded
        return ERR_DETACHING, recorded

    def check(self, typ, model, node, device_id) -> str:
        mid, err = machine_id_from_annotations(self.f, node)
        if err:
            return err
        body, err = self.machine_info(mid)
        if err:
            return err
        return _o.fabric_check_resource("cm", body, typ, model, device_id)

    def resources(self) -> Tuple[List[Dict[str, str]], str]:
        out: List[Dict[str, str]] = []
        for n in self.f.node_names():
            mid, err = machine_id_from_annotations(self.f, n)
            if err:
                return [], err
            body, err = self.machine_info(mid)
            if err:
                return [], err
            out += _o.fabric_get_resources("cm", body, n, mid)
        return out, ""


def select_adapter(env: Dict[str, str]) -> Tuple[str, str]:
    """composableresource_adapter.go:39-72 -> (kind, err)."""
    drt = env.get("DEVICE_RESOURCE_TYPE", "")
    if drt not in ("DEVICE_PLUGIN", "DRA"):
        return "", "the env variable DEVICE_RESOURCE_TYPE has an invalid value: '%s'" % drt
    p = env.get("CDI_PROVIDER_TYPE", "")
    if p == "SUNFISH":
        return "sunfish", ""
    if p == "FTI_CDI":
        if env.get("FTI_CDI_CLUSTER_ID", "") == "" and drt == "DEVICE_PLUGIN":
            return "", "The cluster in RKE2 does not support DEVICE_PLUGIN, please use DRA"
        api = env.get("FTI_CDI_API_TYPE", "")
        if api in ("CM", "FM"):
            return api.lower(), ""
        return "", "the env variable FTI_CDI_API_TYPE has an invalid value: '%s'" % api
    return "", "the env variable CDI_PROVIDER_TYPE has an invalid value: '%s'" % p


# --------------------------------------------------------------------------
# time.Parse(time.RFC3339, ...) and the DaemonSet restart rule
# --------------------------------------------------------------------------
_LAYOUT = "2006-01-02T15:04:05Z07:00"


def _tq(s: str) -> str:
    out = '"'
    for b in s.encode("utf-8", "surrogateescape"):
        if b >= 0x80 or b < 0x20:
            out += "\\x%02x" % b
        else:
            if b in (0x22, 0x5C):
                out += "\\"
            out += chr(b)
    return out + '"'


def parse_rfc3339(value: str) -> Tuple[Optional[Tuple[int, int]], str]:
    """((unix seconds, nanoseconds), "") or (None, ParseError text)."""
    import calendar
    v = value

    def cannot(elem_value: str, elem: str):
        return None, "parsing time %s as %s: cannot parse %s as %s" % (_tq(value), _tq(_LAYOUT), _tq(elem_value), _tq(elem))

    def msg(m: str):
        return None, "parsing time %s%s" % (_tq(value), m)

    def num(s: str, fixed: bool):
        if not s[:1].isascii() or not s[:1].isdigit():
            return None, s
        if not (s[1:2].isascii() and s[1:2].isdigit()):
            return (None, s) if fixed else (int(s[0]), s[1:])
        return int(s[:2]), s[2:]

    if len(v.encode("utf-8", "surrogateescape")) < 4 or not re.match(r"[0-9]", v):
        return cannot(v, "2006")
    if not re.match(r"[0-9]{4}", v):
        return cannot(v, "2006")
    year, v = int(v[:4]), v[4:]
    fields = []
    for lit, elem, fixed in (("-", "01", True), ("-", "02", True), ("T", "15", False), (":", "04", True), (":", "05", True)):
        if v[:1] != lit:
            return cannot(v, lit)
        v = v[1:]
        x, rest = num(v, fixed)
        if x is None:
            return cannot(v, elem)
        v = rest
        if elem == "01" and not 1 <= x <= 12:
            return msg(": month out of range")
        if elem == "15" and x >= 24:
            return msg(": hour out of range")
        if elem == "04" and x >= 60:
            return msg(": minute out of range")
        if elem == "05" and x >= 60:
            return msg(": second out of range")
        fields.append(x)
    month, day, hour, minute, sec = fields
    nsec = 0
    m = re.match(r"[.,]([0-9]+)", v)
    if m:
        digits = m.group(1)
        nsec = int((digits[:9]).ljust(9, "0"))
        v = v[m.end():]
    offset = 0
    if v[:1] == "Z":
        v = v[1:]
    else:
        hold = v
        if len(v.encode("utf-8", "surrogateescape")) < 6 or v[3:4] != ":":
            return cannot(hold, "Z07:00")
        sign, hh, mm, v = v[0], v[1:3], v[4:6], v[6:]
        hr, _ = num(hh, True)
        mn = None
        if hr is not None:
            mn, _ = num(mm, True)
        ok = hr is not None and mn is not None and sign in "+-"
        rng = ""
        if (hr or 0) > 24:
            rng = "time zone offset hour"
        if (mn or 0) > 60:
            rng = "time zone offset minute"
        if rng:
            return msg(": " + rng + " out of range")
        if not ok:
            return cannot(hold, "Z07:00")
        offset = (hr * 60 + mn) * 60 * (-1 if sign == "-" else 1)
    if v != "":
        return msg(": extra text: " + _tq(v))
    leap = year % 4 == 0 and (year % 100 != 0 or year % 400 == 0)
    if day < 1 or day > (29 if (month == 2 and leap) else calendar.monthrange(2001, month)[1]):
        return msg(": day out of range")
    days = _days_from_civil(year, month, day)
    return (days * 86400 + hour * 3600 + minute * 60 + sec - offset, nsec), ""


def _days_from_civil(y: int, m: int, d: int) -> int:
    y -= m <= 2
    era = (y if y >= 0 else y - 399) // 400
    yoe = y - era * 400
    doy = (153 * (m + (-3 if m > 2 else 9)) + 2) // 5 + d - 1
    doe = yoe * 365 + yoe // 4 - yoe // 100 + doy
    return era * 146097 + doe - 719468


def restart_daemonset(ns: str, name: str, ds: Dict, now: Tuple[int, int]) -> Tuple[bool, str]:
    """internal/utils/nodes.go:35-76 after a successful Get -> (restart issued?, err)."""
    desired = int(ds.get("desired", 0))
    if desired == 0:
        return False, ""
    if (int(ds.get("ready", 0)) < desired or int(ds.get("current", 0)) < desired or int(ds.get("unavailable", 0)) > 0 or
            int(ds.get("misscheduled", 0)) > 0):
        return False, ""
    if isinstance(ds.get("restarted_at"), str):
        t, err = parse_rfc3339(ds["restarted_at"])
        if err:
            return False, "failed to parse restartedAt annotation for DaemonSet %s/%s: '%s'" % (ns, name, err)
        since = (now[0] - t[0]) * 10 ** 9 + (now[1] - t[1])
        if since <= 10 * 10 ** 9:
            return False, ""
    return True, ""


class SunfishClient:
    """internal/cdi/sunfish/client.go:76-146 (no reference test exercises it: parity of the body is derived)."""
    MODELS = ("Tesla-V100-PCIE-16GB", "NVIDIA-A100-PCIE-40GB", "NVIDIA-A100-80GB-PCIe")

    def __init__(self, fabric: Fabric):
        self.f = fabric

    def _patch(self, node: str, model: str, count: int) -> str:
        known = model in self.MODELS
        body = _o.emit_sunfish(node, count if known else 0, "GPU" if known else "", model if known else "")
        st, _, terr = self.f.do("PATCH", "redfish/v1/Systems/System", "", body)
        if terr:
            return terr
        if st not in (200, 204):
            return "http returned code %d" % st
        return ""

    def add(self, node: str, model: str) -> Tuple[str, str, str]:
        return "", "", self._patch(node, model, 1)

    def remove(self, node: str, model: str) -> str:
        return self._patch(node, model, 0)
