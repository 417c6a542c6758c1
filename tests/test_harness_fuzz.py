"""The JSON-driven entry points must survive any request: wrong types in every position, missing keys, huge or empty
values.  (The parsers behind them run under ASan/UBSan in tests/test_host_sanitizers.py; this drives the real
shared library end to end, where a crash would take the operator down.)"""
import json
import random

KEYS = ["name", "spec", "status", "labels", "deleting", "device_resource_type", "probe", "provider", "enumeration", "resource_slices",
        "daemonsets", "daemonset_errors", "now", "env", "fabric", "cluster", "load_check", "drain", "enumeration_after_remove",
        "resource_slices_after_remove", "driver_pod_missing", "http", "objects", "nodes", "metal3machines", "baremetalhosts",
        "annotations", "pods", "exec", "needle", "escape", "literal", "cluster_policy", "driver_enabled", "state", "device_id",
        "cdi_device_id", "type", "model", "target_node", "force_detach", "stdout", "stderr", "exec_err", "method", "path", "status",
        "body", "token_error", "taints", "devices", "attributes", "uuid", "containers", "namespace", "node", "fd_scan", "remove",
        "DEVICE_RESOURCE_TYPE", "CDI_PROVIDER_TYPE", "FTI_CDI_API_TYPE", "FTI_CDI_TENANT_ID", "FTI_CDI_CLUSTER_ID", "op", "proc_root"]
ATOMS = [None, True, False, 0, -1, 1 << 40, 1.5, "", "x", "Attaching", "Detaching", "Online", "DRA", "DEVICE_PLUGIN", "FTI_CDI", "CM", "FM",
         "SUNFISH", "GPU-1", "worker-0", "2025-01-01T00:00:00Z", "error", "{}", "[", " 😀", "a" * 5000, "drain", "check_no_gpu_loads"]


def rand_value(rng, depth):
    r = rng.random()
    if depth <= 0 or r < 0.35:
        return rng.choice(ATOMS)
    if r < 0.55:
        return [rand_value(rng, depth - 1) for _ in range(rng.randrange(0, 4))]
    return {rng.choice(KEYS): rand_value(rng, depth - 1) for _ in range(rng.randrange(0, 7))}


def test_reconcile_attach_and_friends_never_crash(cro):
    rng = random.Random(31337)
    ok = errors = 0
    for _ in range(3000):
        req = rand_value(rng, 5)
        if not isinstance(req, dict):
            req = {"status": req}
        if rng.random() < 0.5:
            req.setdefault("status", {})
            if isinstance(req["status"], dict):
                req["status"]["state"] = rng.choice(["", "Attaching", "Online", "Detaching", "Deleting", "bogus"])
        req["probe"] = False                    # no device on the CPU box; the probe path has its own tests
        for fn in (lambda r: cro.reconcile_attach(None, r), cro.fabric_list_devices):
            try:
                out = fn(req)
                assert isinstance(out, dict)
                ok += 1
            except cro.ProbeError:
                errors += 1
        if isinstance(req.get("cluster"), dict) or rng.random() < 0.2:
            local = dict(req, op=rng.choice(["check_no_gpu_loads", "run_nvidia_smi", "check_gpu_visible", "drain", "nope"]),
                         proc_root="/nonexistent", allow_mutation=False)
            try:
                cro.local_node_op(None, local)
                ok += 1
            except cro.ProbeError:
                errors += 1
    assert ok > 1000 and errors > 0
    # and raw bytes that are not even JSON
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    n = ctypes.c_size_t(0)
    for raw in (b"", b"{", b"null", b"[]", b"\xff\xfe", b'{"status":' * 2000, json.dumps({"fabric": {"http": [[]] * 1000}}).encode()):
        assert cro.lib.cro_reconcile_attach(None, raw, buf, 1 << 16, ctypes.byref(n)) in (cro.OK, cro.ERR_PARSE, cro.ERR_INVALID_ARG, cro.ERR_BUFFER_SMALL)
