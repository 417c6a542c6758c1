"""composable-resource-operator_b200 — B200-native post-attach probe + spec path.

Thin ctypes binding of ``libcroprobe.so`` (the C ABI in ``include/croprobe.h``),
the same entry points a Go host binds with cgo.  The package
name is not a Python identifier; import it with::

    import importlib
    cro = importlib.import_module("composable-resource-operator_b200")

There is no CPU fallback: if the shared library is missing the import raises,
and ``ProbeContext`` raises ``ProbeError`` when no CUDA device is usable.  The
host-side mirror of the reference interface (parse / decide / emit / attach
step) lives in the library too (``csrc/identity.cpp``, ``csrc/reconcile.cpp``);
this module only marshals arguments.

Reference slots (paths relative to the reference tree):
  enumerate / parse     internal/utils/gpus.go:878-919, 921-962, 1014-1089
  visibility decision   internal/utils/gpus.go:54-86
  attach step           internal/controller/composableresource_controller.go:200-287
  wire structs          internal/cdi/fti/fm/api/*.go, fti/cm/client.go:62-79,
                        sunfish/client.go:48-61, api/v1alpha1/*_types.go
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Dict, List, Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcroprobe.so")

ABI_VERSION = 2
MAX_DEVICES = 16

OK = 0
ERR_INVALID_ARG, ERR_ABI_MISMATCH, ERR_NO_DEVICE, ERR_CUDA, ERR_OOM = -1, -2, -3, -4, -5
ERR_CHECKSUM, ERR_BUFFER_SMALL, ERR_NCCL, ERR_DEADLINE, ERR_UNSUPPORTED = -6, -7, -8, -9, -10
ERR_PARSE, ERR_EXEC, ERR_P2P, ERR_INTERNAL = -11, -12, -13, -14

F_SKIP_COPY, F_SKIP_P2P, F_SKIP_NCCL, F_NO_NVML, F_VERIFY_COPY, F_LAZY_ALLOC, F_DEGRADE_ON_OOM, F_SKIP_P2P_WRITE = 1, 2, 4, 8, 16, 32, 64, 128
F_TEST_INJECT = 256
READ_AUTO, READ_LDG, READ_TMA, READ_LDG256 = 0, 1, 2, 3
COPY_AUTO, COPY_LDG, COPY_TMA, COPY_TMA_FUSED = 0, 1, 2, 3
DEV_IN_PROCESS, DEV_NEEDS_HELPER = 1, 2
FAIL_NONE, FAIL_EXPECT, FAIL_COPY_SRC, FAIL_READ, FAIL_P2P_READ, FAIL_P2P_PUSH, FAIL_P2P_CHASE, FAIL_STALE = range(8)


class ProbeError(RuntimeError):
    def __init__(self, code: int, msg: str = "") -> None:
        self.code = code
        super().__init__("croprobe error %d (%s)%s" % (code, strerror(code), (": " + msg) if msg else ""))


class Opts(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_uint32), ("flags", ctypes.c_uint32),
        ("sweep_bytes", ctypes.c_uint64), ("p2p_bytes", ctypes.c_uint64), ("seed_base", ctypes.c_uint64),
        ("read_sweeps", ctypes.c_uint32), ("copy_sweeps", ctypes.c_uint32), ("latency_hops", ctypes.c_uint32),
        ("read_variant", ctypes.c_uint32), ("copy_variant", ctypes.c_uint32), ("deadline_ms", ctypes.c_int32),
        ("n_devices", ctypes.c_int32), ("devices", ctypes.c_int32 * MAX_DEVICES),
        ("rank_base", ctypes.c_uint32), ("world_override", ctypes.c_uint32),
        ("test_inject_after", ctypes.c_uint32), ("reserved0", ctypes.c_uint32),
        ("test_inject_word", ctypes.c_uint64), ("test_inject_mask", ctypes.c_uint64),
    ]


class DevInfo(ctypes.Structure):
    _fields_ = [
        ("cuda_ordinal", ctypes.c_int32), ("device_minor", ctypes.c_int32),
        ("gpu_uuid", ctypes.c_char * 48), ("pci_bus_id", ctypes.c_char * 24), ("name", ctypes.c_char * 64),
        ("hbm_bytes_total", ctypes.c_uint64), ("sm_count", ctypes.c_uint32),
        ("cc_major", ctypes.c_uint32), ("cc_minor", ctypes.c_uint32), ("identity_source", ctypes.c_uint32),
        ("flags", ctypes.c_uint32), ("dev_index", ctypes.c_int32), ("reserved", ctypes.c_uint32 * 2),
    ]


class ProbeResult(ctypes.Structure):
    """cro_probe_result (ABI 2): written on the device by the finalize kernel, 512 bytes."""
    _fields_ = [
        ("abi_version", ctypes.c_uint32), ("status", ctypes.c_int32),
        ("cuda_ordinal", ctypes.c_int32), ("device_minor", ctypes.c_int32),
        ("gpu_uuid", ctypes.c_char * 48), ("pci_bus_id", ctypes.c_char * 24),
        ("hbm_bytes_total", ctypes.c_uint64), ("sweep_bytes", ctypes.c_uint64), ("seed", ctypes.c_uint64),
        ("checksum_xor", ctypes.c_uint64), ("checksum_sum", ctypes.c_uint64),
        ("fill_ns", ctypes.c_uint64), ("read_best_ns", ctypes.c_uint64), ("read_median_ns", ctypes.c_uint64),
        ("copy_best_ns", ctypes.c_uint64), ("copy_median_ns", ctypes.c_uint64),
        ("sm_count", ctypes.c_uint32), ("sm_clock_mhz", ctypes.c_uint32), ("mem_clock_mhz", ctypes.c_uint32),
        ("ecc_errors", ctypes.c_uint32),
        ("p2p_read_ns", ctypes.c_uint64 * 8), ("p2p_checksum_xor", ctypes.c_uint64 * 8),
        ("p2p_latency_ns_x16", ctypes.c_uint32 * 8), ("p2p_access", ctypes.c_uint8 * 8),
        ("p2p_bytes", ctypes.c_uint64), ("expect_xor", ctypes.c_uint64), ("expect_sum", ctypes.c_uint64),
        ("expect_wsum", ctypes.c_uint64), ("checksum_wsum", ctypes.c_uint64),
        ("copy_checksum_xor", ctypes.c_uint64), ("copy_checksum_sum", ctypes.c_uint64), ("copy_checksum_wsum", ctypes.c_uint64),
        ("total_ns", ctypes.c_uint64), ("p2p_write_ns", ctypes.c_uint64 * 8),
        ("nonce", ctypes.c_uint32), ("rank", ctypes.c_uint8), ("world", ctypes.c_uint8),
        ("read_variant", ctypes.c_uint8), ("copy_variant", ctypes.c_uint8),
        ("read_sweeps", ctypes.c_uint8), ("copy_sweeps", ctypes.c_uint8), ("copy_verified", ctypes.c_uint8),
        ("fail_code", ctypes.c_uint8), ("fail_index", ctypes.c_uint8), ("p2p_ok", ctypes.c_uint8),
        ("reserved8", ctypes.c_uint8 * 2), ("t_start_ns", ctypes.c_uint64),
    ]

    @property
    def checksum(self) -> Tuple[int, int, int]:
        return (self.checksum_xor, self.checksum_sum, self.checksum_wsum)

    @property
    def expect(self) -> Tuple[int, int, int]:
        return (self.expect_xor, self.expect_sum, self.expect_wsum)

    @property
    def copy_checksum(self) -> Tuple[int, int, int]:
        return (self.copy_checksum_xor, self.copy_checksum_sum, self.copy_checksum_wsum)


class SweepResult(ctypes.Structure):
    _fields_ = [("bytes", ctypes.c_uint64), ("ns", ctypes.c_uint64), ("checksum_xor", ctypes.c_uint64),
                ("checksum_sum", ctypes.c_uint64), ("variant", ctypes.c_uint32), ("launches", ctypes.c_uint32),
                ("checksum_wsum", ctypes.c_uint64), ("timer_ns", ctypes.c_uint64)]

    @property
    def checksum(self) -> Tuple[int, int, int]:
        return (self.checksum_xor, self.checksum_sum, self.checksum_wsum)


class SweepTime(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("index", ctypes.c_uint32), ("bytes", ctypes.c_uint64),
                ("event_ns", ctypes.c_uint64), ("timer_ns", ctypes.c_uint64)]


class P2PDetail(ctypes.Structure):
    _fields_ = [("read_ns", ctypes.c_uint64), ("push_ns", ctypes.c_uint64), ("reread_ns", ctypes.c_uint64),
                ("read_xor", ctypes.c_uint64), ("read_sum", ctypes.c_uint64), ("read_wsum", ctypes.c_uint64),
                ("landed_xor", ctypes.c_uint64), ("landed_sum", ctypes.c_uint64), ("landed_wsum", ctypes.c_uint64),
                ("expect_xor", ctypes.c_uint64), ("expect_sum", ctypes.c_uint64), ("expect_wsum", ctypes.c_uint64),
                ("chase_ns", ctypes.c_uint64),
                ("chase_end", ctypes.c_uint32), ("chase_expect", ctypes.c_uint32), ("hops", ctypes.c_uint32), ("access", ctypes.c_uint32)]


class FullBoxTime(ctypes.Structure):
    _fields_ = [("enqueue_ns", ctypes.c_uint64), ("wall_ns", ctypes.c_uint64), ("hbm_ns", ctypes.c_uint64),
                ("p2p_ns", ctypes.c_uint64), ("chase_ns", ctypes.c_uint64), ("gather_ns", ctypes.c_uint64),
                ("rounds", ctypes.c_uint32), ("host_syncs", ctypes.c_uint32), ("gather", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32)]


GATHER_HOST, GATHER_NCCL, GATHER_DEGRADED = 0, 1, 2


assert ctypes.sizeof(ProbeResult) == 512, ctypes.sizeof(ProbeResult)

# Every symbol include/croprobe.h declares; tests check the library exports all of them.
EXPORTS = [
    "cro_probe_init", "cro_probe_destroy", "cro_device_count", "cro_enumerate", "cro_emit_csv",
    "cro_parse_gpu_csv", "cro_parse_proc_csv", "cro_proc_information_to_line", "cro_check_gpu_visible",
    "cro_normalize", "cro_probe_device", "cro_probe_all", "cro_result_device_ptr", "cro_hbm_fill",
    "cro_hbm_read_checksum", "cro_hbm_copy", "cro_hbm_read_checksum_dst", "cro_hbm_expected_checksum",
    "cro_inject_fault", "cro_read_words", "cro_hbm_read_loop", "cro_hbm_copy_loop", "cro_hbm_fill_loop",
    "cro_device_seed", "cro_launch_count", "cro_emit_status_json", "cro_emit_scalar_status_json",
    "cro_emit_fm_scale_up", "cro_emit_fm_scale_down", "cro_emit_cm_scale_up", "cro_emit_cm_scale_down",
    "cro_emit_sunfish_request", "cro_emit_probe_annotations_json", "cro_fm_parse_scale_up_response",
    "cro_reconcile_attach", "cro_strerror", "cro_last_error", "cro_version", "cro_cm_check_adding_resources",
    "cro_sim_create", "cro_sim_destroy", "cro_sim_apply", "cro_sim_delete", "cro_sim_plant", "cro_sim_run",
    "cro_sim_reconcile_request", "cro_sim_dump", "cro_probe_begin", "cro_probe_end",
    "cro_check_no_gpu_loads", "cro_check_gpu_drain_status", "cro_check_device_file_scan",
    "cro_scan_device_file_holders", "cro_sim_reconcile_resource", "cro_sim_sync_upstream",
    "cro_fabric_check_resource", "cro_fabric_get_resources", "cro_fabric_list_devices",
    "cro_local_node_op", "cro_scan_cmdline_for", "cro_token_from_reply",
    "cro_selftest_exception_barrier", "cro_probe_sweep_times", "cro_p2p_detail_get", "cro_fullbox_times",
    "cro_chase_end", "cro_validate_env", "cro_node_inventory", "cro_probe_uuid", "cro_set_latency_hops", "cro_local_exec", "cro_metrics_text", "cro_describe_wire_type",
]


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libcroprobe.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C composable-resource-operator_b200/csrc`.  There is no CPU fallback for the probe path."
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c, sz, psz = ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)
    vp, i32, u32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
    out = [c, sz, psz]
    sig = {
        "cro_probe_init": (i32, [ctypes.POINTER(Opts), ctypes.POINTER(vp)]),
        "cro_probe_destroy": (None, [vp]),
        "cro_device_count": (i32, [vp, ctypes.POINTER(i32)]),
        "cro_enumerate": (i32, [vp, ctypes.POINTER(DevInfo), i32, ctypes.POINTER(i32)]),
        "cro_emit_csv": (i32, [ctypes.POINTER(DevInfo), i32, c] + out),
        "cro_parse_gpu_csv": (i32, [c, c, c, c] + out),
        "cro_parse_proc_csv": (i32, [c, c, c, c] + out),
        "cro_proc_information_to_line": (i32, [c] + out),
        "cro_check_gpu_visible": (i32, [ctypes.POINTER(DevInfo), i32, c, ctypes.POINTER(i32)]),
        "cro_normalize": (i32, [i32, c] + out),
        "cro_probe_device": (i32, [vp, i32, ctypes.POINTER(ProbeResult)]),
        "cro_probe_all": (i32, [vp, ctypes.POINTER(ProbeResult), i32, ctypes.POINTER(i32)]),
        "cro_probe_begin": (i32, [vp, i32]),
        "cro_probe_end": (i32, [vp, i32, ctypes.POINTER(ProbeResult)]),
        "cro_result_device_ptr": (i32, [vp, i32, ctypes.POINTER(u64)]),
        "cro_hbm_fill": (i32, [vp, i32, ctypes.POINTER(SweepResult)]),
        "cro_hbm_fill_loop": (i32, [vp, i32, u32, ctypes.POINTER(SweepResult)]),
        "cro_hbm_read_checksum": (i32, [vp, i32, u32, ctypes.POINTER(SweepResult)]),
        "cro_hbm_read_checksum_dst": (i32, [vp, i32, u32, ctypes.POINTER(SweepResult)]),
        "cro_hbm_read_loop": (i32, [vp, i32, u32, u32, ctypes.POINTER(SweepResult)]),
        "cro_hbm_copy": (i32, [vp, i32, u32, ctypes.POINTER(SweepResult)]),
        "cro_hbm_copy_loop": (i32, [vp, i32, u32, u32, ctypes.POINTER(SweepResult)]),
        "cro_hbm_expected_checksum": (i32, [vp, i32, ctypes.POINTER(SweepResult)]),
        "cro_inject_fault": (i32, [vp, i32, u64, u64]),
        "cro_read_words": (i32, [vp, i32, u64, u64, ctypes.POINTER(u64)]),
        "cro_device_seed": (i32, [vp, i32, ctypes.POINTER(u64)]),
        "cro_probe_sweep_times": (i32, [vp, i32, ctypes.POINTER(SweepTime), i32, ctypes.POINTER(i32)]),
        "cro_p2p_detail_get": (i32, [vp, i32, i32, ctypes.POINTER(P2PDetail)]),
        "cro_fullbox_times": (i32, [vp, ctypes.POINTER(FullBoxTime)]),
        "cro_chase_end": (i32, [i32, i32, u32, ctypes.POINTER(u32)]),
        "cro_set_latency_hops": (i32, [vp, u32]),
        "cro_metrics_text": (i32, [vp] + out),
        "cro_validate_env": (i32, [c, c, c, sz]),
        "cro_node_inventory": (i32, [c, ctypes.POINTER(DevInfo), i32, ctypes.POINTER(DevInfo), i32, ctypes.POINTER(i32)]),
        "cro_probe_uuid": (i32, [vp, c, ctypes.POINTER(ProbeResult)]),
        "cro_launch_count": (u64, [vp]),
        "cro_emit_status_json": (i32, [c, c, c, c] + out),
        "cro_emit_scalar_status_json": (i32, [c, c, c, c, c] + out),
        "cro_emit_fm_scale_up": (i32, [c, c, c, c] + out),
        "cro_emit_fm_scale_down": (i32, [c, c, c, c] + out),
        "cro_emit_cm_scale_up": (i32, [c, i32] + out),
        "cro_emit_cm_scale_down": (i32, [c, i32, c] + out),
        "cro_emit_sunfish_request": (i32, [c, ctypes.c_longlong, c, c] + out),
        "cro_emit_probe_annotations_json": (i32, [ctypes.POINTER(ProbeResult)] + out),
        "cro_fm_parse_scale_up_response": (i32, [c, c, c, c, c, sz, c, sz, c, sz]),
        "cro_reconcile_attach": (i32, [vp, c] + out),
        "cro_cm_check_adding_resources": (i32, [c, c, c, c, c, sz, ctypes.POINTER(i32), c, sz, c, sz, c, sz]),
        "cro_sim_create": (i32, [vp, c, ctypes.POINTER(vp)]),
        "cro_sim_destroy": (None, [vp]),
        "cro_sim_apply": (i32, [vp, c, c, sz]),
        "cro_sim_plant": (i32, [vp, c, c, sz]),
        "cro_sim_delete": (i32, [vp, c]),
        "cro_sim_run": (i32, [vp, ctypes.c_longlong] + out),
        "cro_sim_reconcile_request": (i32, [vp, c, c, sz]),
        "cro_sim_dump": (i32, [vp] + out),
        "cro_fabric_check_resource": (i32, [c, c, c, c, c, c, sz]),
        "cro_fabric_get_resources": (i32, [c, c, c, c] + out),
        "cro_fabric_list_devices": (i32, [c] + out),
        "cro_token_from_reply": (i32, [c] + out),
        "cro_selftest_exception_barrier": (i32, [i32]),
        "cro_local_node_op": (i32, [vp, c] + out),
        "cro_local_exec": (i32, [c] + out),
        "cro_describe_wire_type": (i32, [c] + out),
        "cro_scan_cmdline_for": (i32, [c, c, ctypes.POINTER(i32)]),
        "cro_sim_reconcile_resource": (i32, [vp, c, c, sz]),
        "cro_sim_sync_upstream": (i32, [vp, c, ctypes.c_longlong, c, sz]),
        "cro_check_no_gpu_loads": (i32, [c, c, c, c, c, c, i32, c, sz]),
        "cro_check_gpu_drain_status": (i32, [c, c, c, c, c, ctypes.POINTER(i32), c, sz]),
        "cro_check_device_file_scan": (i32, [c, c, c, i32, c, sz]),
        "cro_scan_device_file_holders": (i32, [c, c, i32] + out),
        "cro_strerror": (c, [i32]),
        "cro_last_error": (i32, [vp, c, sz]),
        "cro_version": (c, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError == a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def strerror(code: int) -> str:
    return lib.cro_strerror(code).decode()


def version() -> str:
    return lib.cro_version().decode()


def _b(s) -> Optional[bytes]:
    if s is None:
        return None
    return s if isinstance(s, bytes) else s.encode("utf-8", "surrogateescape")


def _text_call(fn, *args, cap: int = 1 << 16) -> Tuple[int, bytes]:
    buf = ctypes.create_string_buffer(cap)
    n = ctypes.c_size_t(0)
    rc = fn(*args, buf, cap, ctypes.byref(n))
    if rc == ERR_BUFFER_SMALL and n.value + 1 > cap:
        return _text_call(fn, *args, cap=n.value + 1)
    return rc, buf.raw[: n.value]


def _text(fn, *args) -> str:
    rc, raw = _text_call(fn, *args)
    if rc != OK:
        raise ProbeError(rc, raw.decode("utf-8", "replace"))
    return raw.decode("utf-8", "surrogateescape")


# ---- host-side mirror of the reference text path (no GPU needed) -------------------
def getGPUInfoFromNvidiaSmiOutput(std_out: str, std_err: str, exec_err: Optional[str], queryArgs: str) -> Tuple[int, str]:
    """Parse rule of getGPUInfoFromNvidiaPod (internal/utils/gpus.go:896-916).
    Returns (code, text): Go JSON of the []map[string]string, or the error text."""
    rc, raw = _text_call(lib.cro_parse_gpu_csv, _b(std_out), _b(std_err), _b(exec_err), _b(queryArgs))
    return rc, raw.decode("utf-8", "surrogateescape")


def getGPUInfoFromProcOutput(std_out: str, std_err: str, exec_err: Optional[str], queryArgs: str) -> Tuple[int, str]:
    """Parse rule of getGPUInfoFromProcInCroNodeAgentPod (internal/utils/gpus.go:1045-1089)."""
    rc, raw = _text_call(lib.cro_parse_proc_csv, _b(std_out), _b(std_err), _b(exec_err), _b(queryArgs))
    return rc, raw.decode("utf-8", "surrogateescape")


def proc_information_to_line(text: str) -> str:
    return _text(lib.cro_proc_information_to_line, _b(text))


def normalize(kind: int, s: str) -> str:
    return _text(lib.cro_normalize, kind, _b(s))


def emit_csv(devs: List[DevInfo], query: str) -> str:
    arr = (DevInfo * max(1, len(devs)))(*devs)
    return _text(lib.cro_emit_csv, arr, len(devs), _b(query))


def CheckGPUVisible(devs: List[DevInfo], device_id: str) -> bool:
    """internal/utils/gpus.go:73-84 over an enumerated device list."""
    arr = (DevInfo * max(1, len(devs)))(*devs)
    v = ctypes.c_int(0)
    rc = lib.cro_check_gpu_visible(arr, len(devs), _b(device_id), ctypes.byref(v))
    if rc != OK:
        raise ProbeError(rc)
    return bool(v.value)


def emit_status_json(state: str, error: str = "", device_id: str = "", cdi_device_id: str = "") -> str:
    return _text(lib.cro_emit_status_json, _b(state), _b(error), _b(device_id), _b(cdi_device_id))


def emit_scalar_status_json(state: str, device_id: str = "", cdi_device_id: str = "", node_name: str = "",
                            error: str = "") -> str:
    return _text(lib.cro_emit_scalar_status_json, _b(state), _b(device_id), _b(cdi_device_id), _b(node_name), _b(error))


def emit_fm_scale_up(tenant: str, mach: str, res_type: str, model: str) -> str:
    return _text(lib.cro_emit_fm_scale_up, _b(tenant), _b(mach), _b(res_type), _b(model))


def emit_fm_scale_down(tenant: str, mach: str, res_type: str, res_uuid: str) -> str:
    return _text(lib.cro_emit_fm_scale_down, _b(tenant), _b(mach), _b(res_type), _b(res_uuid))


def emit_cm_scale_up(spec_uuid: str, device_count: int) -> str:
    return _text(lib.cro_emit_cm_scale_up, _b(spec_uuid), device_count)


def emit_cm_scale_down(spec_uuid: str, device_count: int, device_id: str) -> str:
    return _text(lib.cro_emit_cm_scale_down, _b(spec_uuid), device_count, _b(device_id))


def emit_sunfish_request(name: str, count: int, proc_type: str, model: str) -> str:
    return _text(lib.cro_emit_sunfish_request, _b(name), count, _b(proc_type), _b(model))


def emit_probe_annotations_json(r: ProbeResult) -> str:
    return _text(lib.cro_emit_probe_annotations_json, ctypes.byref(r))


def fm_parse_scale_up_response(body: str, name: str, res_type: str, model: str) -> Tuple[str, str, str]:
    """(deviceID, CDIDeviceID, err) per internal/cdi/fti/fm/client.go:184-213."""
    dev, cdi, err = (ctypes.create_string_buffer(256) for _ in range(3))
    err = ctypes.create_string_buffer(1024)
    rc = lib.cro_fm_parse_scale_up_response(_b(body), _b(name), _b(res_type), _b(model), dev, 256, cdi, 256, err, 1024)
    if rc != OK:
        return "", "", err.value.decode()
    return dev.value.decode(), cdi.value.decode(), ""


def cm_check_adding_resources(machine_body: str, existing_device_ids: List[str], res_type: str, model: str):
    """(specUUID, deviceCount, deviceID, CDIDeviceID, err) per internal/cdi/fti/cm/client.go:432-459."""
    spec, dev, cdi = (ctypes.create_string_buffer(256) for _ in range(3))
    err = ctypes.create_string_buffer(1024)
    n = ctypes.c_int(0)
    lib.cro_cm_check_adding_resources(_b(machine_body), _b("\n".join(existing_device_ids)), _b(res_type), _b(model),
                                      spec, 256, ctypes.byref(n), dev, 256, cdi, 256, err, 1024)
    return spec.value.decode(), n.value, dev.value.decode(), cdi.value.decode(), err.value.decode()


def reconcile_attach(ctx: Optional["ProbeContext"], request: Dict) -> Dict:
    """One pass of handleAttachingState (composableresource_controller.go:200-287).
    ``ctx`` may be None when ``request['enumeration']`` supplies the exec output."""
    handle = ctx.handle if ctx is not None else None
    rc, raw = _text_call(lib.cro_reconcile_attach, handle, _b(json.dumps(request)))
    if rc != OK:
        raise ProbeError(rc, raw.decode("utf-8", "replace"))
    out = json.loads(raw.decode("utf-8"))
    out["_raw"] = raw.decode("utf-8")
    return out


# ---- the probe context (needs a B200) ------------------------------------------------
class ProbeContext:
    """Long-lived probe context: resident sweep buffers, streams, events.

    Takes the slot of utils.RunNvidiaSmi + utils.CheckGPUVisible
    (internal/utils/gpus.go:666-689, 54-86)."""

    def __init__(self, sweep_bytes: int = 0, devices: Optional[List[int]] = None, flags: int = 0,
                 read_sweeps: int = 0, copy_sweeps: int = 0, read_variant: int = READ_AUTO,
                 copy_variant: int = COPY_AUTO, p2p_bytes: int = 0, latency_hops: int = 0,
                 deadline_ms: int = 0, seed_base: int = 0, rank_base: int = 0, world: int = 0,
                 inject: Optional[Tuple[int, int, int]] = None) -> None:
        o = Opts()
        o.abi_version = ABI_VERSION
        o.flags = flags
        o.sweep_bytes = sweep_bytes
        o.p2p_bytes = p2p_bytes
        o.seed_base = seed_base
        o.read_sweeps, o.copy_sweeps = read_sweeps, copy_sweeps
        o.latency_hops = latency_hops
        o.read_variant, o.copy_variant = read_variant, copy_variant
        o.deadline_ms = deadline_ms
        o.rank_base, o.world_override = rank_base, world
        if inject is not None:            # (after sweep number, word index, xor mask): fault injection inside the probe
            o.flags |= F_TEST_INJECT
            o.test_inject_after, o.test_inject_word, o.test_inject_mask = inject
        if devices:
            o.n_devices = len(devices)
            for i, d in enumerate(devices):
                o.devices[i] = d
        h = ctypes.c_void_p()
        rc = lib.cro_probe_init(ctypes.byref(o), ctypes.byref(h))
        if rc != OK:
            why = ctypes.create_string_buffer(1024)
            lib.cro_last_error(None, why, 1024)       # the context died with the init: its text is per-thread
            detail = why.value.decode("utf-8", "replace")
            raise ProbeError(rc, "cro_probe_init" + (": " + detail if detail else ""))
        self.handle = h

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib.cro_probe_destroy(self.handle)
            self.handle = None

    def __enter__(self) -> "ProbeContext":
        return self

    def __exit__(self, *a) -> None:
        self.close()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow=()) -> int:
        if rc != OK and rc not in allow:
            buf = ctypes.create_string_buffer(1024)
            lib.cro_last_error(self.handle, buf, 1024)
            raise ProbeError(rc, buf.value.decode("utf-8", "replace"))
        return rc

    def last_error(self) -> str:
        """Text of this thread's most recent failing (or degrading) call on the context."""
        buf = ctypes.create_string_buffer(1024)
        lib.cro_last_error(self.handle, buf, 1024)
        return buf.value.decode("utf-8", "replace")

    def device_count(self) -> int:
        n = ctypes.c_int()
        self._check(lib.cro_device_count(self.handle, ctypes.byref(n)))
        return n.value

    def enumerate(self) -> List[DevInfo]:
        n = ctypes.c_int()
        arr = (DevInfo * MAX_DEVICES)()
        self._check(lib.cro_enumerate(self.handle, arr, MAX_DEVICES, ctypes.byref(n)))
        return [arr[i] for i in range(n.value)]

    def own_devices(self) -> List[DevInfo]:
        """The devices this context probes in process, by dev_index (enumerate() lists the whole NODE, fresh)."""
        mine = [d for d in self.enumerate() if d.flags & DEV_IN_PROCESS]
        return sorted(mine, key=lambda d: d.dev_index)

    def seed(self, dev: int = 0) -> int:
        s = ctypes.c_uint64()
        self._check(lib.cro_device_seed(self.handle, dev, ctypes.byref(s)))
        return s.value

    def probe_device(self, dev: int = 0, allow_checksum_error: bool = False) -> ProbeResult:
        r = ProbeResult()
        self._check(lib.cro_probe_device(self.handle, dev, ctypes.byref(r)),
                    allow=(ERR_CHECKSUM,) if allow_checksum_error else ())
        return r

    def probe_begin(self, dev: int = 0) -> None:
        self._check(lib.cro_probe_begin(self.handle, dev))

    def probe_end(self, dev: int = 0) -> ProbeResult:
        r = ProbeResult()
        self._check(lib.cro_probe_end(self.handle, dev, ctypes.byref(r)))
        return r

    def probe_all(self) -> List[ProbeResult]:
        arr = (ProbeResult * MAX_DEVICES)()
        n = ctypes.c_int()
        self._check(lib.cro_probe_all(self.handle, arr, MAX_DEVICES, ctypes.byref(n)))
        return [arr[i] for i in range(n.value)]

    def result_device_ptr(self, dev: int = 0) -> int:
        p = ctypes.c_uint64()
        self._check(lib.cro_result_device_ptr(self.handle, dev, ctypes.byref(p)))
        return p.value

    def _sweep(self, fn, *args) -> SweepResult:
        r = SweepResult()
        self._check(fn(self.handle, *args, ctypes.byref(r)))
        return r

    def hbm_fill(self, dev: int = 0, iters: int = 1) -> SweepResult:
        return self._sweep(lib.cro_hbm_fill_loop, dev, iters)

    def hbm_read_checksum(self, dev: int = 0, variant: int = READ_AUTO, iters: int = 1, dst: bool = False) -> SweepResult:
        if dst:
            return self._sweep(lib.cro_hbm_read_checksum_dst, dev, variant)
        return self._sweep(lib.cro_hbm_read_loop, dev, variant, iters)

    def hbm_copy(self, dev: int = 0, variant: int = COPY_AUTO, iters: int = 1) -> SweepResult:
        return self._sweep(lib.cro_hbm_copy_loop, dev, variant, iters)

    def hbm_expected_checksum(self, dev: int = 0) -> SweepResult:
        return self._sweep(lib.cro_hbm_expected_checksum, dev)

    def inject_fault(self, dev: int, word_index: int, mask: int) -> None:
        self._check(lib.cro_inject_fault(self.handle, dev, word_index, mask))

    def read_words(self, dev: int, first: int, n: int) -> List[int]:
        arr = (ctypes.c_uint64 * n)()
        self._check(lib.cro_read_words(self.handle, dev, first, n, arr))
        return list(arr)

    def launch_count(self) -> int:
        return int(lib.cro_launch_count(self.handle))

    def sweep_times(self, dev: int = 0) -> List[SweepTime]:
        """Per-sweep CUDA-event and %globaltimer times of the device's last probe (fill, copies, reads)."""
        arr = (SweepTime * 64)()
        n = ctypes.c_int()
        self._check(lib.cro_probe_sweep_times(self.handle, dev, arr, 64, ctypes.byref(n)))
        return [arr[i] for i in range(n.value)]

    def p2p_detail(self, dev: int, peer: int) -> P2PDetail:
        d = P2PDetail()
        self._check(lib.cro_p2p_detail_get(self.handle, dev, peer, ctypes.byref(d)))
        return d

    def metrics_text(self) -> str:
        """Prometheus text exposition of the context's counters and per-GPU gauges."""
        rc, raw = _text_call(lib.cro_metrics_text, self.handle)
        self._check(rc)
        return raw.decode()

    def set_latency_hops(self, hops: int) -> None:
        self._check(lib.cro_set_latency_hops(self.handle, hops))

    def fullbox_times(self) -> FullBoxTime:
        t = FullBoxTime()
        self._check(lib.cro_fullbox_times(self.handle, ctypes.byref(t)))
        return t


def node_inventory(proc_root: Optional[str], in_process: List[DevInfo]) -> List[DevInfo]:
    """What cro_enumerate answers for a context managing `in_process` on a node whose /proc is at proc_root:
    the driver's registry is re-read, devices attached later are flagged DEV_NEEDS_HELPER, removed ones are dropped."""
    arr = (DevInfo * max(1, len(in_process)))(*in_process)
    out = (DevInfo * 64)()
    n = ctypes.c_int()
    rc = lib.cro_node_inventory(_b(proc_root), arr, len(in_process), out, 64, ctypes.byref(n))
    if rc != OK:
        raise ProbeError(rc, "cro_node_inventory")
    return [out[i] for i in range(n.value)]


def probe_uuid(ctx: Optional["ProbeContext"], uuid: str) -> ProbeResult:
    """cro_probe_uuid: in-process probe, or the helper process for a GPU attached after init (ctx may be None)."""
    r = ProbeResult()
    rc = lib.cro_probe_uuid(ctx.handle if ctx is not None else None, _b(uuid), ctypes.byref(r))
    if rc not in (OK, ERR_CHECKSUM):
        buf = ctypes.create_string_buffer(1024)
        lib.cro_last_error(ctx.handle if ctx is not None else None, buf, 1024)
        raise ProbeError(rc, buf.value.decode("utf-8", "replace"))
    return r


def chase_end(minor_src: int, minor_dst: int, hops: int) -> int:
    """Slot reached after `hops` steps of the latency permutation of the directed pair (host arithmetic)."""
    e = ctypes.c_uint32()
    rc = lib.cro_chase_end(minor_src, minor_dst, hops, ctypes.byref(e))
    if rc != OK:
        raise ProbeError(rc, "cro_chase_end")
    return e.value


def validate_env(name: Optional[str] = None, value: Optional[str] = None) -> str:
    """"" when the CRO_* knob (or, with no name, the process environment) is legal, else the reference-style
    sentence "the env variable X has an invalid value: 'v'" (composableresource_adapter.go:44)."""
    err = ctypes.create_string_buffer(512)
    rc = lib.cro_validate_env(_b(name), _b(value), err, 512)
    return "" if rc == OK else err.value.decode("utf-8", "replace")


def fabric_check_resource(kind: str, machine_body: str, res_type: str, model: str, device_id: str) -> str:
    """CdiProvider.CheckResource decision (fm/client.go:314-359, cm/client.go:262-304); "" = healthy."""
    err = ctypes.create_string_buffer(2048)
    lib.cro_fabric_check_resource(_b(kind), _b(machine_body), _b(res_type), _b(model), _b(device_id), err, 2048)
    return err.value.decode("utf-8", "surrogateescape")


def fabric_get_resources(kind: str, machine_body: str, node_name: str, machine_uuid: str) -> List[Dict]:
    """CdiProvider.GetResources decode for one node (fm/client.go:385-410, cm/client.go:335-343)."""
    return json.loads(_text(lib.cro_fabric_get_resources, _b(kind), _b(machine_body), _b(node_name), _b(machine_uuid)))


def fabric_list_devices(request: Dict) -> Dict:
    """CdiProvider.GetResources of the FM / CM client over a scripted fabric (what the UpstreamSyncer
    tick reads: upstreamsyncer_controller.go:77-84).  request = {"env": {...}, "fabric": {...}}."""
    return json.loads(_text(lib.cro_fabric_list_devices, _b(json.dumps(request))))


def token_from_reply(reply: Dict) -> Dict:
    """What fti.CachedToken.Token makes of the id_manager's answer (fti/token.go:96-175):
    reply = {"secret_error","transport_error","status","body"} -> {"error", "expiry"}."""
    return json.loads(_text(lib.cro_token_from_reply, _b(json.dumps(reply))))


def local_node_op(ctx: Optional["ProbeContext"], request: Dict) -> Dict:
    """One node-side operation of internal/utils/gpus.go run locally (scans native, read-only commands spawned,
    mutating ones only with allow_mutation).  See cro_local_node_op in include/croprobe.h."""
    return json.loads(_text(lib.cro_local_node_op, ctx.handle if ctx is not None else None, _b(json.dumps(request))))


def describe_wire_type(name: str) -> Dict:
    """The reply struct a fabric decoder walks, as the library describes it (declaration order)."""
    return json.loads(_text(lib.cro_describe_wire_type, _b(name)))


def local_exec(argv: List[str], allow_mutation: bool = False, exec_deadline_ms: int = 0, native_nvml: bool = True,
               nvml_lib: str = "") -> Dict:
    """One command through the node-local executor (read-only allow-list while allow_mutation is false; deadline).
    The detach side's nvidia-smi invocations (compute apps, drain -q/-m/-r, -pm) are answered through NVML in this
    process when it is there ("how": "native"); nvml_lib names another libnvidia-ml (tests load a stand-in)."""
    req = {"argv": argv, "allow_mutation": allow_mutation, "native_nvml": native_nvml}
    if nvml_lib:
        req["nvml_lib"] = nvml_lib
    if exec_deadline_ms:
        req["exec_deadline_ms"] = exec_deadline_ms
    return json.loads(_text(lib.cro_local_exec, _b(json.dumps(req))))


def scan_cmdline_for(proc_root: str, needle: str) -> bool:
    found = ctypes.c_int(0)
    rc = lib.cro_scan_cmdline_for(_b(proc_root), _b(needle), ctypes.byref(found))
    if rc != OK:
        raise ProbeError(rc, "cro_scan_cmdline_for")
    return bool(found.value)


def CheckNoGPULoadsFromOutput(std_out: str, std_err: str, exec_err: Optional[str], pod_name: str, node_name: str,
                              target_uuid: Optional[str], driver_enabled: bool) -> str:
    """utils.CheckNoGPULoads parse + decision (internal/utils/gpus.go:145-186); returns the error text ("" = nil)."""
    err = ctypes.create_string_buffer(4096)
    lib.cro_check_no_gpu_loads(_b(std_out), _b(std_err), _b(exec_err), _b(pod_name), _b(node_name), _b(target_uuid),
                               int(driver_enabled), err, 4096)
    return err.value.decode("utf-8", "surrogateescape")


def checkGPUDrainStatusFromOutput(std_out: str, std_err: str, exec_err: Optional[str], node_name: str,
                                  bus_id: str) -> Tuple[bool, str]:
    """checkGPUDrainStatus (internal/utils/gpus.go:964-1012); returns (draining, error text)."""
    err = ctypes.create_string_buffer(4096)
    d = ctypes.c_int(0)
    lib.cro_check_gpu_drain_status(_b(std_out), _b(std_err), _b(exec_err), _b(node_name), _b(bus_id), ctypes.byref(d), err, 4096)
    return bool(d.value), err.value.decode("utf-8", "surrogateescape")


def CheckDeviceFileScanResult(std_out: str, std_err: str, exec_err: Optional[str], rke2: bool = False) -> str:
    err = ctypes.create_string_buffer(4096)
    lib.cro_check_device_file_scan(_b(std_out), _b(std_err), _b(exec_err), int(rke2), err, 4096)
    return err.value.decode("utf-8", "surrogateescape")


def scan_device_file_holders(target: str, proc_root: Optional[str] = None, rke2: bool = False) -> str:
    """Native fd scan (replaces the shell scripts at internal/utils/gpus.go:236-260, 441-457)."""
    return _text(lib.cro_scan_device_file_holders, _b(proc_root), _b(target), int(rke2))


class Cluster:
    """In-memory API server + both reconcilers (cro_sim_*): the caller of the hot path.

    Mirrors ComposabilityRequestReconciler / ComposableResourceReconciler
    (internal/controller/*.go); drives the storm / churn configs."""

    def __init__(self, config: Dict, ctx: Optional[ProbeContext] = None) -> None:
        h = ctypes.c_void_p()
        rc = lib.cro_sim_create(ctx.handle if ctx is not None else None, _b(json.dumps(config)), ctypes.byref(h))
        if rc != OK:
            raise ProbeError(rc, "cro_sim_create")
        self.handle = h
        self._ctx = ctx   # keep the probe context alive

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib.cro_sim_destroy(self.handle)
            self.handle = None

    def __enter__(self) -> "Cluster":
        return self

    def __exit__(self, *a) -> None:
        self.close()

    def _err_call(self, fn, arg: str) -> str:
        err = ctypes.create_string_buffer(1024)
        fn(self.handle, _b(arg), err, 1024)
        return err.value.decode("utf-8", "replace")

    def apply(self, name: str, resource: Dict) -> str:
        """kubectl apply of a ComposabilityRequest; returns "" or the admission error."""
        return self._err_call(lib.cro_sim_apply, json.dumps({"name": name, "resource": resource}))

    def plant(self, obj: Dict) -> str:
        return self._err_call(lib.cro_sim_plant, json.dumps(obj))

    def delete(self, name: str) -> bool:
        return lib.cro_sim_delete(self.handle, _b(name)) == OK

    def reconcile_request(self, name: str) -> str:
        """One Reconcile of the request controller; returns the reconcile error ("" = nil)."""
        return self._err_call(lib.cro_sim_reconcile_request, name)

    def reconcile_resource(self, name: str) -> str:
        """One Reconcile of the ComposableResource controller; returns the reconcile error ("" = nil)."""
        return self._err_call(lib.cro_sim_reconcile_resource, name)

    def sync_upstream(self, devices: List[Dict], now_s: int) -> str:
        """One UpstreamSyncer tick (upstreamsyncer_controller.go:77-136) at time now_s."""
        err = ctypes.create_string_buffer(1024)
        lib.cro_sim_sync_upstream(self.handle, _b(json.dumps(devices)), now_s, err, 1024)
        return err.value.decode("utf-8", "replace")

    def run(self, max_reconciles: int = 0) -> Dict:
        rc, raw = _text_call(lib.cro_sim_run, self.handle, max_reconciles, cap=1 << 16)
        if rc != OK:
            raise ProbeError(rc)
        return json.loads(raw.decode())

    def dump(self) -> Dict:
        rc, raw = _text_call(lib.cro_sim_dump, self.handle, cap=1 << 22)
        if rc != OK:
            raise ProbeError(rc)
        return json.loads(raw.decode())
