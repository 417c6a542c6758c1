"""CPU oracle, Python side — TEST INFRASTRUCTURE ONLY.

Two things live here:

* ``PyOracle``: a readable pure-Python / numpy restatement of the reference's
  enumerate -> parse -> decide -> emit path (file:line cited per function,
  paths relative to the reference tree) and of the probe pattern's closed form.
* ``COracle``: ctypes binding of ``liboracle.so`` (``cro_oracle.c``), the C
  restatement used at full size and as the timed CPU baseline.

The two are written independently and are cross-checked against each other and
against ``tests/golden/reference_kats.json`` (strings transcribed from the
reference's own Ginkgo tests).  The reference is Go and cannot be compiled in
this image, so there is no ``oracle/_ref``; see the header of cro_oracle.c for
the parity status.  Product code (``composable-resource-operator_b200/``) never
imports this module.
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import go_decode  # noqa: E402  (same directory; how encoding/json fills the wire structs)

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MASK = (1 << 64) - 1

OK, ERR_EXEC, ERR_PARSE, ERR_UNSUPPORTED = 0, -12, -11, -10


# --------------------------------------------------------------------------
# probe pattern (new work)
# --------------------------------------------------------------------------
def pattern_word(seed: int, i: int) -> int:
    z = (seed + i + 0x9E3779B97F4A7C15) & MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


def pattern_words_np(seed: int, first: int, n_words: int) -> np.ndarray:
    """Vectorised ``pattern_word`` over [first, first+n_words) (uint64 wraps)."""
    with np.errstate(over="ignore"):
        i = np.arange(n_words, dtype=np.uint64) + np.uint64(first & MASK)
        z = i + np.uint64(seed & MASK) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def checksum_np(seed: int, first: int, n_words: int, chunk: int = 1 << 22, pos0: Optional[int] = None) -> Tuple[int, int, int]:
    """(XOR, wrapping sum, position-weighted sum: sum of w * (2*pos + 1)) of pattern words [first, first+n_words),
    word first+i at position pos0+i (default: pos0 = first)."""
    x, s, ws = 0, 0, 0
    done = 0
    pos0 = first if pos0 is None else pos0
    while done < n_words:
        n = min(chunk, n_words - done)
        w = pattern_words_np(seed, first + done, n)
        x ^= int(np.bitwise_xor.reduce(w))
        with np.errstate(over="ignore"):
            s = (s + int(np.add.reduce(w, dtype=np.uint64))) & MASK
            pos = np.arange(n, dtype=np.uint64) + np.uint64((pos0 + done) & MASK)
            ws = (ws + int(np.add.reduce(w * (np.uint64(2) * pos + np.uint64(1)), dtype=np.uint64))) & MASK
        done += n
    return x, s, ws


NONCE_STRIDE = 0xD1B54A32D192ED03


def probe_seed(seed_base: int, minor: int, nonce: int) -> int:
    """Seed of probe number `nonce` on a device (nonce 0 = SURVEY.md §8d's seed_base | minor)."""
    return ((seed_base | minor) + nonce * NONCE_STRIDE) & MASK


# --------------------------------------------------------------------------
# Go string helpers
# --------------------------------------------------------------------------
_GO_SPACE = set("\t\n\v\f\r \x85\xa0\u1680\u2028\u2029\u202f\u205f\u3000") | {chr(c) for c in range(0x2000, 0x200B)}


def go_trim_space(s: str) -> str:
    b, e = 0, len(s)
    while b < e and s[b] in _GO_SPACE:
        b += 1
    while e > b and s[e - 1] in _GO_SPACE:
        e -= 1
    return s[b:e]


def go_json_string(s: str) -> str:
    """encoding/json appendString with escapeHTML=true (Go 1.24)."""
    out = ['"']
    for ch in s:
        c = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif c < 0x20 or ch in "<>&":
            out.append("\\u%04x" % c)
        elif c in (0x2028, 0x2029):
            out.append("\\u%04x" % c)
        elif 0xDC80 <= c <= 0xDCFF:  # surrogateescape'd invalid byte -> U+FFFD
            out.append("\\ufffd")
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def go_json_bytes(raw: bytes) -> bytes:
    """Same, for a byte string that may hold invalid UTF-8 (one \\ufffd per bad byte)."""
    return go_json_string(raw.decode("utf-8", "surrogateescape")).encode("utf-8", "surrogateescape")


def go_marshal_string_map(m: Dict[str, str]) -> str:
    keys = sorted(m.keys(), key=lambda k: k.encode("utf-8", "surrogateescape"))
    return "{" + ",".join(go_json_string(k) + ":" + go_json_string(m[k]) for k in keys) + "}"


# --------------------------------------------------------------------------
# parse rules
# --------------------------------------------------------------------------
@dataclass
class ParseResult:
    code: int = OK
    infos: Optional[List[Dict[str, str]]] = None  # None == Go nil slice
    error: str = ""

    def to_json(self) -> str:
        if self.infos is None:
            return "null"
        return "[" + ",".join(go_marshal_string_map(m) for m in self.infos) + "]"


def _exec_error(std_out: str, std_err: str, exec_err: Optional[str]) -> str:
    return "get gpu info command failed: err: '%s', stderr: '%s', stdout: '%s'" % (
        exec_err if exec_err is not None else "<nil>", std_err, std_out)


def parse_gpu_csv(std_out: str, std_err: str, exec_err: Optional[str], query: str) -> ParseResult:
    """internal/utils/gpus.go:880,896-916 (getGPUInfoFromNvidiaPod; twin at :923,939-959)."""
    field_names = query.split(",")
    if go_trim_space(std_out) == "No devices were found":  # :896 — before the error test
        return ParseResult(OK, [])
    if std_err != "" or exec_err is not None:  # :899
        return ParseResult(ERR_EXEC, None, _exec_error(std_out, std_err, exec_err))
    infos: Optional[List[Dict[str, str]]] = None
    for line in go_trim_space(std_out).split("\n"):
        if line == "":
            continue
        parts = line.split(",")
        info: Dict[str, str] = {}
        for i, name in enumerate(field_names):
            if i >= len(parts):  # :913 parts[i] unguarded -> Go panics
                return ParseResult(ERR_PARSE, None,
                                   "runtime error: index out of range [%d] with length %d" % (i, len(parts)))
            info[name] = go_trim_space(parts[i])
        infos = (infos or []) + [info]
    return ParseResult(OK, infos)


def parse_proc_csv(std_out: str, std_err: str, exec_err: Optional[str], query: str) -> ParseResult:
    """internal/utils/gpus.go:1045-1089 (getGPUInfoFromProcInCroNodeAgentPod)."""
    field_names = query.split(",")
    if std_err != "" or exec_err is not None:
        return ParseResult(ERR_EXEC, None, _exec_error(std_out, std_err, exec_err))
    trimmed = go_trim_space(std_out)
    if trimmed == "":
        return ParseResult(OK, [])
    infos: Optional[List[Dict[str, str]]] = None
    for line in trimmed.split("\n"):
        if line == "":
            continue
        parts = line.split(",")
        if len(parts) < 3:
            return ParseResult(ERR_PARSE, None, "unexpected GPU information format: '%s'" % line)
        values = {"device_minor": go_trim_space(parts[0]), "gpu_uuid": go_trim_space(parts[1]),
                  "pci.bus_id": go_trim_space(parts[2])}
        info: Dict[str, str] = {}
        for f in field_names:
            name = go_trim_space(f)
            if name not in values:
                return ParseResult(ERR_UNSUPPORTED, None, "unsupported field '%s' requested in queryArgs" % name)
            info[name] = values[name]
        infos = (infos or []) + [info]
    return ParseResult(OK, infos)


def proc_information_to_line(text: str) -> str:
    """The awk lines + printf at internal/utils/gpus.go:1030-1034."""
    def third(key: str) -> str:
        for line in text.split("\n"):
            if line.startswith(key):
                f = line.replace("\t", " ").split()
                return f[2] if len(f) >= 3 else ""
        return ""
    minor, uuid, bus = third("Device Minor:"), third("GPU UUID:"), third("Bus Location:")
    if not (minor and uuid and bus):
        return ""
    return "%s,%s,%s\n" % (minor, uuid, bus)


def check_gpu_visible(std_out: str, std_err: str, exec_err: Optional[str], device_id: str) -> Tuple[bool, str]:
    """internal/utils/gpus.go:73-84 (DEVICE_PLUGIN branch)."""
    r = parse_gpu_csv(std_out, std_err, exec_err, "gpu_uuid")
    if r.code != OK:
        return False, r.error
    for info in r.infos or []:
        if info["gpu_uuid"] == device_id:
            return True, ""
    return False, ""


def normalize(kind: int, s: str) -> str:
    """internal/utils/gpus.go:218 (0), :326 (1), :406,567 (2), :238 (3), :480 (4)."""
    def up(t: str) -> str:
        return "".join(chr(ord(c) - 32) if "a" <= c <= "z" else c for c in t)

    def low(t: str) -> str:
        return "".join(chr(ord(c) + 32) if "A" <= c <= "Z" else c for c in t)
    if kind == 0:
        return up(go_trim_space(s))
    if kind == 1:
        return low(go_trim_space(s))
    if kind == 2:
        t = up(go_trim_space(s))
        return t[4:] if t.startswith("0000") else t
    if kind == 3:
        return "/dev/nvidia" + s
    if kind == 4:
        return "/run/nvidia/driver/dev/nvidia" + s
    raise ValueError(kind)


# --------------------------------------------------------------------------
# emitters (json.Marshal of the wire structs)
# --------------------------------------------------------------------------
def emit_status(state: str, error: str = "", device_id: str = "", cdi_device_id: str = "") -> str:
    """api/v1alpha1/composableresource_types.go:36-41."""
    parts = ['"state":' + go_json_string(state)]
    if error:
        parts.append('"error":' + go_json_string(error))
    if device_id:
        parts.append('"device_id":' + go_json_string(device_id))
    if cdi_device_id:
        parts.append('"cdi_device_id":' + go_json_string(cdi_device_id))
    return "{" + ",".join(parts) + "}"


def emit_scalar_status(state: str, device_id: str = "", cdi_device_id: str = "", node_name: str = "",
                       error: str = "") -> str:
    """api/v1alpha1/composabilityrequest_types.go:74-80."""
    parts = ['"state":' + go_json_string(state)]
    for k, v in (("device_id", device_id), ("cdi_device_id", cdi_device_id), ("node_name", node_name),
                 ("error", error)):
        if v:
            parts.append('"%s":%s' % (k, go_json_string(v)))
    return "{" + ",".join(parts) + "}"


def emit_fm_scale_up(tenant: str, mach: str, res_type: str, model: str) -> str:
    """internal/cdi/fti/fm/api/scale_up.go:19-41, common.go:21-29; fti/fm/client.go:115-144."""
    cond = '{"column":"model","operator":"eq","value":%s}' % go_json_string(model)
    spec = '{"res_type":%s,"res_spec":{"condition":[%s]},"res_num":1}' % (go_json_string(res_type), cond)
    return '{"tenants":{"tenant_uuid":%s,"machines":[{"mach_uuid":%s,"resources":[{"res_specs":[%s]}]}]}}' % (
        go_json_string(tenant), go_json_string(mach), spec)


def emit_fm_scale_down(tenant: str, mach: str, res_type: str, res_uuid: str) -> str:
    """internal/cdi/fti/fm/api/scale_down.go:19-41; fti/fm/client.go:247-271."""
    spec = '{"res_type":%s,"res_uuid":%s,"res_num":1}' % (go_json_string(res_type), go_json_string(res_uuid))
    return '{"tenants":{"tenant_uuid":%s,"machines":[{"mach_uuid":%s,"resources":[{"res_specs":[%s]}]}]}}' % (
        go_json_string(tenant), go_json_string(mach), spec)


def emit_cm_scale_up(spec_uuid: str, device_count: int) -> str:
    """internal/cdi/fti/cm/client.go:62-69."""
    return '{"increase_resource_count":{"spec_uuid":%s,"device_count":%d}}' % (go_json_string(spec_uuid), device_count)


def emit_cm_scale_down(spec_uuid: str, device_count: int, device_id: str) -> str:
    """internal/cdi/fti/cm/client.go:71-79."""
    return '{"remove_resources":{"spec_uuid":%s,"device_count":%d,"devices":[%s]}}' % (
        go_json_string(spec_uuid), device_count, go_json_string(device_id))


def emit_sunfish(name: str, count: int, proc_type: str, model: str) -> str:
    """internal/cdi/sunfish/client.go:48-61."""
    return ('{"Name":%s,"Processors":{"Members":[{"@Redfish.RequestCount":%d,"ProcessorType":%s,"Model":%s}]}}'
            % (go_json_string(name), count, go_json_string(proc_type), go_json_string(model)))


# --------------------------------------------------------------------------
# Go's struct decoding of JSON objects
# --------------------------------------------------------------------------
def _fold(s: str) -> str:
    """bytes.EqualFold's normal form for keys compared with ASCII field names (U+017F -> s, U+212A -> k)."""
    return "".join("s" if ch == "\u017f" else "k" if ch == "\u212a" else (ch.lower() if "A" <= ch <= "Z" else ch) for ch in s)


class GoObj(dict):
    """A JSON object as encoding/json sees it when it fills a struct: keys come in input order and each key that
    names a field — exactly or case-folded — overwrites what an earlier one stored, so the LAST such key wins."""

    @classmethod
    def from_pairs(cls, pairs):
        o = cls(pairs)
        o.pairs = list(pairs)
        return o

    def _field(self, key):
        hit, found = None, False
        f = _fold(key)
        for k, v in getattr(self, "pairs", list(self.items())):
            if k == key or _fold(k) == f:
                hit, found = v, True
        return hit, found

    def get(self, key, default=None):
        v, found = self._field(key)
        return v if found else default

    def __getitem__(self, key):
        v, found = self._field(key)
        if not found:
            raise KeyError(key)
        return v

    def __contains__(self, key):
        return self._field(key)[1]


def go_loads(text: str):
    return json.loads(text, object_pairs_hook=GoObj.from_pairs)


def go_struct_loads(text: str, go_type: str):
    """What a zero `go_type` holds after json.Unmarshal(text, &v): go_decode.decode_as over the wire-struct
    descriptions (null leaves a field alone, repeated members merge, repeated slices decode over earlier elements)."""
    return go_decode.decode_as(text, go_decode.TYPES[go_type])


# --------------------------------------------------------------------------
# FM gate + attach step
# --------------------------------------------------------------------------
def fm_scale_up_response_to_ids(body: str, name: str, spec_type: str, spec_model: str) -> Tuple[str, str, str]:
    """internal/cdi/fti/fm/client.go:184-213.  Returns (deviceID, CDIDeviceID, err)."""
    data = go_struct_loads(body, "api.ScaleUpResponse")
    machines = (data.get("data") or {}).get("machines") or []
    if machines and (machines[0].get("resources") or []) and machines[0]["resources"][0].get("res_type", "") == spec_type:
        res = machines[0]["resources"][0]
        for c in ((res.get("res_spec") or {}).get("condition") or []):
            if c.get("column") == "model" and c.get("operator") == "eq" and c.get("value") == spec_model:
                op = res.get("res_op_status", "")
                if op == "":
                    return "", "", "runtime error: slice bounds out of range [:1] with length 0"
                if op[:1] in ("0", "1"):
                    return res.get("res_serial_num", ""), res.get("res_uuid", ""), ""
                if op[:1] == "2":
                    return "", "", "the FM attached device called by %s is in Critical state in FM" % name
                return "", "", "the FM attached device called by %s is in unknown state '%s' in FM" % (name, op)
    return "", "", "can not find the added gpu when using FM to add gpu"


def _op_status(op: str, device_id: str, where: str) -> str:
    if op == "":
        return "runtime error: slice bounds out of range [:1] with length 0"
    if op[:1] == "0":
        return ""
    if op[:1] == "1":
        return "the target gpu '%s' is showing a Warning status in %s" % (device_id, where)
    if op[:1] == "2":
        return "the target gpu '%s' is showing a Critical status in %s" % (device_id, where)
    return "the target gpu '%s' has unknown status '%s' in %s" % (device_id, op, where)


def fabric_check_resource(kind: str, body: str, spec_type: str, spec_model: str, device_id: str) -> str:
    """FM: internal/cdi/fti/fm/client.go:314-359.  CM: internal/cdi/fti/cm/client.go:262-304."""
    data = go_struct_loads(body, "api.GetMachineResponse" if kind == "fm" else "api.MachineData")
    if kind == "fm":
        machines = (data.get("data") or {}).get("machines") or []
        if not machines:
            return "runtime error: index out of range [0] with length 0"
        for r in machines[0].get("resources") or []:
            if r.get("res_type", "") != spec_type:
                continue
            for c in ((r.get("res_spec") or {}).get("condition") or []):
                if c.get("column") != "model" or c.get("operator") != "eq" or c.get("value") != spec_model:
                    continue
                if r.get("res_serial_num", "") == device_id:
                    return _op_status(r.get("res_op_status", ""), device_id, "FM")
    else:
        specs = ((((data.get("data") or {}).get("cluster") or {}).get("machine") or {}).get("resspecs")) or []
        for s in specs:
            if s.get("type", "") != spec_type:
                continue
            for c in ((((s.get("selector") or {}).get("expression") or {}).get("conditions")) or []):
                if c.get("column") != "model" or c.get("operator") != "eq" or c.get("value") != spec_model:
                    continue
                for d in s.get("devices") or []:
                    if d.get("device_id", "") == device_id:
                        return _op_status((d.get("detail") or {}).get("res_op_status", ""), device_id, "CM")
    return "the target device '%s' cannot be found in CDI system" % device_id


def fabric_get_resources(kind: str, body: str, node: str, machine_uuid: str) -> List[Dict[str, str]]:
    """FM: internal/cdi/fti/fm/client.go:385-410.  CM: internal/cdi/fti/cm/client.go:335-343."""
    data = go_struct_loads(body, "api.GetMachineResponse" if kind == "fm" else "api.MachineData")
    out = []
    if kind == "fm":
        machines = (data.get("data") or {}).get("machines") or []
        if not machines:
            return out
        for r in machines[0].get("resources") or []:
            if r.get("res_type", "") != "gpu":
                continue
            model = ""
            for c in ((r.get("res_spec") or {}).get("condition") or []):
                if c.get("column") == "model" and c.get("operator") == "eq":
                    model = c.get("value", "")
                    break
            out.append({"node_name": node, "machine_uuid": machine_uuid, "device_type": "gpu", "model": model,
                        "device_id": r.get("res_serial_num", ""), "cdi_device_id": r.get("res_uuid", "")})
    else:
        specs = ((((data.get("data") or {}).get("cluster") or {}).get("machine") or {}).get("resspecs")) or []
        for s in specs:
            if s.get("type", "") != "gpu":
                continue
            for d in s.get("devices") or []:
                out.append({"node_name": node, "machine_uuid": machine_uuid, "device_type": "gpu", "model": "",
                            "device_id": d.get("device_id", ""), "cdi_device_id": (d.get("detail") or {}).get("res_uuid", "")})
    return out


def check_no_gpu_loads(std_out: str, std_err: str, exec_err: Optional[str], pod_name: str, node_name: str,
                       target_uuid: Optional[str], driver_enabled: bool) -> str:
    """internal/utils/gpus.go:145-186 (parse + decision of CheckNoGPULoads).  Returns the error text."""
    if go_trim_space(std_out) == "No devices were found":
        return ""
    if std_err != "" or exec_err is not None:
        return "run nvidia-smi in pod '%s' to check gpu loads failed: '%s', stderr: '%s', stdout: '%s'" % (
            pod_name, exec_err if exec_err is not None else "<nil>", std_err, std_out)
    apps = []
    for line in go_trim_space(std_out).split("\n"):
        if line == "":
            continue
        parts = line.split(",")
        if len(parts) < 2:
            return "runtime error: index out of range [1] with length %d" % len(parts)
        apps.append((go_trim_space(parts[0]), go_trim_space(parts[1])))
    listed = "[" + " ".join("GPUUUID: '%s', ProcessName: '%s'" % a for a in apps) + "]"   # %v + String() :50-52
    if not driver_enabled:
        if target_uuid is None:
            return "runtime error: invalid memory address or nil pointer dereference"
        if any(a[0] == target_uuid for a in apps):
            return "found gpu load on gpu '%s': %s" % (target_uuid, listed)
        return ""
    if apps:
        return "found gpu loads on node '%s': '%s'" % (node_name, listed)
    return ""


def check_gpu_drain_status(std_out: str, std_err: str, exec_err: Optional[str], node_name: str, bus_id: str):
    """internal/utils/gpus.go:964-1012.  Returns (draining, error text)."""
    def low(t: str) -> str:
        return "".join(chr(ord(c) + 32) if "A" <= c <= "Z" else c for c in t)
    bus = go_trim_space(bus_id)
    if bus == "":
        return False, "target GPU bus ID is empty"
    if exec_err is not None or std_err != "":
        return False, "check gpu drain status command failed: '%s', stderr: '%s', stdout: '%s'" % (
            exec_err if exec_err is not None else "<nil>", std_err, std_out)
    trimmed = go_trim_space(std_out)
    if trimmed == "":
        return False, "nvidia-smi drain query returned empty output (node=%s, busID=%s)" % (node_name, bus)
    for line in trimmed.split("\n"):
        lower = low(go_trim_space(line))
        if "drain" not in lower:
            continue
        idx = lower.find(":")
        if idx >= 0:
            status = go_trim_space(lower[idx + 1:]).strip(".")
            if "not draining" in status:
                return False, ""
            if "draining" in status:
                return True, ""
    return False, "nvidia-smi drain query did not contain recognizable drain state (node=%s, busID=%s, raw=%s)" % (
        node_name, bus, trimmed)


def check_device_file_scan(std_out: str, std_err: str, exec_err: Optional[str], rke2: bool = False) -> str:
    """internal/utils/gpus.go:468-473 / :629-634 (OCP), :286-291 (RKE2)."""
    e = exec_err if exec_err is not None else "<nil>"
    if rke2:
        if exec_err is not None or std_err != "":
            return "deatch command 'check /dev/nvidiaX' failed: '%s', stderr: '%s', stdout: '%s'" % (e, std_err, std_out)
        if std_out != "":
            return "check /dev/nvidiaX command failed: /dev/nvidiaX is in use by one or more processes: %s" % std_out
        return ""
    if std_err != "" or exec_err is not None:
        return "check /dev/nvidiaX command failed: '%s', stderr: '%s'" % (e, std_err)
    if std_out != "":
        return "check /dev/nvidiaX command failed: there is a process %s occupied the nvidiaX file" % std_out
    return ""


def cm_check_adding_resources(machine_body: str, existing_device_ids: List[str], spec_type: str, spec_model: str):
    """internal/cdi/fti/cm/client.go:432-459 (checkAddingResources), :485-499 (isSpecMatch),
    :501-509 (findAvailableDevice).  Returns (specUUID, deviceCount, deviceID, CDIDeviceID, err)."""
    data = go_struct_loads(machine_body, "api.MachineData")
    specs = ((((data.get("data") or {}).get("cluster") or {}).get("machine") or {}).get("resspecs")) or []
    for spec in specs:
        if spec.get("type", "") != spec_type:
            continue
        conds = (((spec.get("selector") or {}).get("expression") or {}).get("conditions")) or []
        if not any(c.get("column") == "model" and c.get("operator") == "eq" and c.get("value") == spec_model for c in conds):
            continue
        for dev in (spec.get("devices") or []):
            if dev.get("device_id", "") in existing_device_ids:
                continue
            res = (dev.get("detail") or {}).get("res_uuid", "")
            if dev.get("status") == "ADD_COMPLETE":
                return "", 0, dev.get("device_id", ""), res, ""
            if dev.get("status") == "ADD_FAILED":
                return "", 0, dev.get("device_id", ""), res, \
                    "an error occurred with the resource in CM: '%s'" % dev.get("status_reason", "")
            break
        return spec.get("spec_uuid", ""), int(spec.get("device_count", 0)), "", "", ""
    return "", 0, "", "", ""


@dataclass
class Status:
    state: str = ""
    error: str = ""
    device_id: str = ""
    cdi_device_id: str = ""

    def to_json(self) -> str:
        return emit_status(self.state, self.error, self.device_id, self.cdi_device_id)


@dataclass
class AttachInput:
    name: str = "test-composable-resource"
    target_node: str = "worker-0"
    deleting: bool = False
    device_resource_type: str = "DEVICE_PLUGIN"
    provider_waiting: bool = False
    provider_error: str = ""
    provider_device_id: str = ""
    provider_cdi_device_id: str = ""
    std_out: str = ""
    std_err: str = ""
    exec_err: Optional[str] = None
    driver_pod_missing: bool = False
    ds_err: Dict[str, str] = field(default_factory=dict)  # "ns/name" -> error
    slice_uuids: Optional[List[str]] = None
    update_fail_after: Optional[int] = None    # Status().Update succeeds this many times, then fails ...
    update_fail_error: str = ""                # ... with this text


def is_go_panic(err: str) -> bool:
    """A Go run-time panic restated as its text.  It unwinds through the handler AND through requeueOnErr — no
    Status().Update on the way — until controller-runtime's Reconcile wrapper recovers it (controller-runtime v0.21.0
    pkg/internal/controller/controller.go: fmt.Errorf("panic: %v [recovered]", r); RecoverPanic defaults to true)."""
    return err.startswith("runtime error: ")


def recovered(err: str) -> str:
    return err if err.startswith("panic: ") else "panic: %s [recovered]" % err


def attach_step(inp: AttachInput, st: Status) -> Tuple[Status, int, str, int]:
    """internal/controller/composableresource_controller.go:200-287 (+ requeueOnErr :423-433).

    Returns (status, requeue_after_s, reconcile_error, n_status_updates) — updates ATTEMPTED; from attempt number
    inp.update_fail_after + 1 on, Status().Update answers inp.update_fail_error and the handler stops where the
    reference stops (`if err := r.Status().Update(...); err != nil { return r.requeueOnErr(...) }`)."""
    st = Status(st.state, st.error, st.device_id, st.cdi_device_id)
    updates = 0
    pod_err = "no Pod with label 'app.kubernetes.io/component=nvidia-driver' found on node %s" % inp.target_node

    def write() -> str:
        nonlocal updates
        updates += 1
        if inp.update_fail_after is not None and updates > inp.update_fail_after:
            return inp.update_fail_error
        return ""

    def requeue_on_err(err: str):
        if is_go_panic(err):                      # never reaches requeueOnErr: no write
            return st, 0, recovered(err), updates
        st.error = err
        write()                                   # :428-430: a failure of this write is only logged
        return st, 0, err, updates

    if inp.deleting:
        if st.device_id == "":
            st.state = "Deleting"
            return st, 0, write(), updates        # :206 `return ctrl.Result{}, r.Status().Update(ctx, resource)`
        if st.error != "":
            st.state = "Detaching"
            return st, 0, write(), updates
    if st.device_id == "":
        if inp.provider_waiting:
            return st, 30, "", updates
        if inp.provider_error:
            return requeue_on_err(inp.provider_error)
        st.error, st.device_id, st.cdi_device_id = "", inp.provider_device_id, inp.provider_cdi_device_id
        e = write()
        if e:
            return requeue_on_err(e)              # :233-235
    if inp.device_resource_type == "DEVICE_PLUGIN":
        for ds in ("nvidia-gpu-operator/nvidia-device-plugin-daemonset", "nvidia-gpu-operator/nvidia-dcgm"):
            if inp.ds_err.get(ds):
                if is_go_panic(inp.ds_err[ds]):
                    return st, 0, recovered(inp.ds_err[ds]), updates
                st.error = inp.ds_err[ds]
                e = write()
                if e:
                    return requeue_on_err(e)
    elif inp.device_resource_type == "DRA":
        e = ""
        if inp.driver_pod_missing:
            e = pod_err
        else:
            r = parse_gpu_csv(inp.std_out, inp.std_err, inp.exec_err, "gpu_uuid")
            if r.code != OK:
                e = r.error
        if e:
            if is_go_panic(e):                    # parts[i] on a short CSV row (gpus.go:912-914): RunNvidiaSmi panics
                return st, 0, recovered(e), updates
            st.error = e
            w = write()
            if w:
                return requeue_on_err(w)
        ds = "nvidia-dra-driver-gpu/nvidia-dra-driver-gpu-kubelet-plugin"
        if inp.ds_err.get(ds):
            if is_go_panic(inp.ds_err[ds]):
                return st, 0, recovered(inp.ds_err[ds]), updates
            st.error = inp.ds_err[ds]
            w = write()
            if w:
                return requeue_on_err(w)
    if inp.device_resource_type == "DRA" and inp.slice_uuids is not None:
        visible, err = (st.device_id in inp.slice_uuids), ""
    elif inp.driver_pod_missing:
        visible, err = False, pod_err
    else:
        visible, err = check_gpu_visible(inp.std_out, inp.std_err, inp.exec_err, st.device_id)
    if err:
        return requeue_on_err(err)
    if visible:
        st.state, st.error = "Online", ""
        return st, 0, write(), updates            # :282
    return st, 30, "", updates


# --------------------------------------------------------------------------
# C oracle binding
# --------------------------------------------------------------------------
class _CStatus(ctypes.Structure):
    _fields_ = [("state", ctypes.c_char * 32), ("error", ctypes.c_char * 1024),
                ("device_id", ctypes.c_char * 128), ("cdi_device_id", ctypes.c_char * 128)]


class _CAttachIn(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("target_node", ctypes.c_char_p), ("deleting", ctypes.c_int),
                ("device_resource_type", ctypes.c_char_p), ("provider_waiting", ctypes.c_int),
                ("provider_error", ctypes.c_char_p), ("provider_device_id", ctypes.c_char_p),
                ("provider_cdi_device_id", ctypes.c_char_p), ("std_out", ctypes.c_char_p),
                ("std_err", ctypes.c_char_p), ("exec_err", ctypes.c_char_p), ("driver_pod_missing", ctypes.c_int),
                ("ds_err", ctypes.c_char_p * 3), ("slice_uuids", ctypes.c_char_p),
                ("update_fail_after", ctypes.c_int), ("update_fail_error", ctypes.c_char_p)]


def build_c_oracle() -> str:
    """Compiles liboracle.so if missing or stale; returns its path."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "cro_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-Wno-format-truncation", "-o", so, src, "-lpthread"])
    return so


def _b(s: Optional[str]) -> Optional[bytes]:
    return None if s is None else s.encode("utf-8", "surrogateescape")


class COracle:
    def __init__(self) -> None:
        self.lib = ctypes.CDLL(build_c_oracle())
        L = self.lib
        u64, p64 = ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)
        L.oracle_pattern_word.restype = u64
        L.oracle_pattern_word.argtypes = [u64, u64]
        L.oracle_checksum.argtypes = [u64, u64, u64, p64, p64]
        L.oracle_checksum3.argtypes = [u64, u64, u64, u64, p64, p64, p64]
        L.oracle_checksum_mt.argtypes = [u64, u64, ctypes.c_int, p64, p64, p64]
        L.oracle_fill.argtypes = [ctypes.c_void_p, u64, u64, u64]
        L.oracle_checksum_buffer.argtypes = [ctypes.c_void_p, u64, p64, p64, p64]
        L.oracle_probe_seed.restype = u64
        L.oracle_probe_seed.argtypes = [u64, ctypes.c_int, u64]
        L.oracle_chase_end.restype = ctypes.c_uint32
        L.oracle_chase_end.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint32]
        for name in ("oracle_parse_gpu_csv", "oracle_parse_proc_csv"):
            getattr(L, name).argtypes = [ctypes.c_char_p] * 4 + [ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_proc_information_to_line.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_check_gpu_visible.argtypes = [ctypes.c_char_p] * 4 + [ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_normalize.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_json_string.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_emit_status.argtypes = [ctypes.c_char_p] * 4 + [ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_emit_scalar_status.argtypes = [ctypes.c_char_p] * 5 + [ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_emit_fm_scale_up.argtypes = [ctypes.c_char_p] * 4 + [ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_emit_fm_scale_down.argtypes = [ctypes.c_char_p] * 4 + [ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_emit_cm_scale_up.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_emit_cm_scale_down.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p,
                                                ctypes.c_size_t]
        L.oracle_emit_sunfish.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_char_p, ctypes.c_char_p,
                                          ctypes.c_char_p, ctypes.c_size_t]
        L.oracle_attach_step.argtypes = [ctypes.POINTER(_CAttachIn), ctypes.POINTER(_CStatus),
                                         ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_int)]
        L.oracle_fm_gate.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]

    # pattern --------------------------------------------------------------
    def pattern_word(self, seed: int, i: int) -> int:
        return int(self.lib.oracle_pattern_word(seed & MASK, i & MASK))

    def checksum(self, seed: int, first: int, n_words: int, threads: int = 1, pos0: Optional[int] = None) -> Tuple[int, int, int]:
        """(xor, sum, position-weighted sum) of pattern words [first, first+n_words)."""
        x, s, w = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        if threads > 1 and first == 0 and pos0 in (None, 0):
            self.lib.oracle_checksum_mt(seed & MASK, n_words, threads, ctypes.byref(x), ctypes.byref(s), ctypes.byref(w))
        else:
            self.lib.oracle_checksum3(seed & MASK, first, n_words, first if pos0 is None else pos0, ctypes.byref(x),
                                      ctypes.byref(s), ctypes.byref(w))
        return x.value, s.value, w.value

    def probe_seed(self, seed_base: int, minor: int, nonce: int) -> int:
        return int(self.lib.oracle_probe_seed(seed_base & MASK, minor, nonce))

    def chase_end(self, minor_src: int, minor_dst: int, hops: int) -> int:
        return int(self.lib.oracle_chase_end(minor_src, minor_dst, hops))

    # text -----------------------------------------------------------------
    def _text(self, fn, *args, cap: int = 1 << 16) -> Tuple[int, str]:
        buf = ctypes.create_string_buffer(cap)
        rc = fn(*args, buf, cap)
        return rc, buf.value.decode("utf-8", "surrogateescape")

    def parse_gpu_csv(self, so, se, ee, q):
        return self._text(self.lib.oracle_parse_gpu_csv, _b(so), _b(se), _b(ee), _b(q))

    def parse_proc_csv(self, so, se, ee, q):
        return self._text(self.lib.oracle_parse_proc_csv, _b(so), _b(se), _b(ee), _b(q))

    def proc_information_to_line(self, text):
        return self._text(self.lib.oracle_proc_information_to_line, _b(text))[1]

    def check_gpu_visible(self, so, se, ee, dev):
        return self._text(self.lib.oracle_check_gpu_visible, _b(so), _b(se), _b(ee), _b(dev))

    def normalize(self, kind, s):
        return self._text(self.lib.oracle_normalize, kind, _b(s))[1]

    def json_string(self, s: str) -> str:
        return self._text(self.lib.oracle_json_string, _b(s))[1]

    def emit_status(self, *a):
        return self._text(self.lib.oracle_emit_status, *[_b(x) for x in a])[1]

    def emit_scalar_status(self, *a):
        return self._text(self.lib.oracle_emit_scalar_status, *[_b(x) for x in a])[1]

    def emit_fm_scale_up(self, *a):
        return self._text(self.lib.oracle_emit_fm_scale_up, *[_b(x) for x in a])[1]

    def emit_fm_scale_down(self, *a):
        return self._text(self.lib.oracle_emit_fm_scale_down, *[_b(x) for x in a])[1]

    def emit_cm_scale_up(self, spec, n):
        return self._text(self.lib.oracle_emit_cm_scale_up, _b(spec), n)[1]

    def emit_cm_scale_down(self, spec, n, dev):
        return self._text(self.lib.oracle_emit_cm_scale_down, _b(spec), n, _b(dev))[1]

    def emit_sunfish(self, name, count, ptype, model):
        return self._text(self.lib.oracle_emit_sunfish, _b(name), count, _b(ptype), _b(model))[1]

    def fm_gate(self, name, op):
        buf = ctypes.create_string_buffer(1024)
        rc = self.lib.oracle_fm_gate(_b(name), _b(op), buf, 1024)
        return rc, buf.value.decode()

    def attach_step(self, inp: AttachInput, st: Status) -> Tuple[Status, int, str, int]:
        cin = _CAttachIn()
        cin.name, cin.target_node = _b(inp.name), _b(inp.target_node)
        cin.deleting = int(inp.deleting)
        cin.device_resource_type = _b(inp.device_resource_type)
        cin.provider_waiting = int(inp.provider_waiting)
        cin.provider_error = _b(inp.provider_error) if inp.provider_error else None
        cin.provider_device_id, cin.provider_cdi_device_id = _b(inp.provider_device_id), _b(inp.provider_cdi_device_id)
        cin.std_out, cin.std_err, cin.exec_err = _b(inp.std_out), _b(inp.std_err), _b(inp.exec_err)
        cin.driver_pod_missing = int(inp.driver_pod_missing)
        names = ("nvidia-gpu-operator/nvidia-device-plugin-daemonset", "nvidia-gpu-operator/nvidia-dcgm",
                 "nvidia-dra-driver-gpu/nvidia-dra-driver-gpu-kubelet-plugin")
        for k, n in enumerate(names):
            cin.ds_err[k] = _b(inp.ds_err[n]) if inp.ds_err.get(n) else None
        cin.slice_uuids = _b("\n".join(inp.slice_uuids)) if inp.slice_uuids is not None else None
        cin.update_fail_after = -1 if inp.update_fail_after is None else inp.update_fail_after
        cin.update_fail_error = _b(inp.update_fail_error)
        cst = _CStatus(_b(st.state), _b(st.error), _b(st.device_id), _b(st.cdi_device_id))
        rq, nu = ctypes.c_int(), ctypes.c_int()
        err = ctypes.create_string_buffer(1024)
        self.lib.oracle_attach_step(ctypes.byref(cin), ctypes.byref(cst), ctypes.byref(rq), err, 1024, ctypes.byref(nu))
        out = Status(cst.state.decode(), cst.error.decode(), cst.device_id.decode(), cst.cdi_device_id.decode())
        return out, rq.value, err.value.decode(), nu.value
