"""Host logic of libcroprobe (parse / decide / emit / attach step) through the
C ABI, against the reference's golden vectors and against the oracle.  No GPU:
none of these entry points touch a device."""
import ctypes
import json
import random

import pytest

from test_oracle_kats import NASTY, _attach_inputs, rand_text


def test_parse_kats(cro, kats):
    for v in kats["parse"]:
        rc, text = cro.getGPUInfoFromNvidiaSmiOutput(v["stdout"], v["stderr"], v["exec_err"], v["query"])
        assert rc == v["code"], v["cite"]
        assert text == (v["json"] if rc == 0 else v["error"]), v["cite"]


def _request(kats, v):
    fx = kats["fixtures"]
    req = {"name": fx["name"], "spec": {"type": "gpu", "model": fx["model"], "target_node": fx["node"]},
           "status": v["status_in"], "deleting": v.get("deleting", False),
           "device_resource_type": v["device_resource_type"], "probe": False}
    if "fm_update" in v:
        req["provider"] = {"fm_response_body": fx["fm_update_body"][v["fm_update"]]}
    else:
        req["provider"] = v.get("provider", {})
    for k in ("enumeration", "driver_pod_missing", "resource_slices", "daemonset_errors"):
        if k in v:
            req[k] = v[k]
    if "enumeration" not in req and not v.get("driver_pod_missing"):
        req["enumeration"] = {"stdout": "", "stderr": ""}
    return req


def test_attach_kats(cro, oracle, kats):
    """The reference's handleAttachingState entries, replayed through cro_reconcile_attach."""
    for v in kats["attach"]:
        out = cro.reconcile_attach(None, _request(kats, v))
        exp = v["expected"]
        assert out["error"] == exp.get("error", ""), (v["cite"], out)
        if "status" in exp:
            e = exp["status"]
            want = oracle.emit_status(e.get("state", ""), e.get("error", ""), e.get("device_id", ""), e.get("cdi_device_id", ""))
            import __graft_entry__ as g
            assert g.json_status(out) == want, (v["cite"], out["_raw"])
            assert out["requeue_after_s"] == exp["requeue_after_s"], v["cite"]


def test_env_error_kat(cro, kats):
    for v in kats["env_errors"]:
        out = cro.reconcile_attach(None, {"status": {"state": "Attaching"}, "device_resource_type": v["device_resource_type"],
                                          "enumeration": {"stdout": ""}})
        assert out["error"] == v["error"]
        assert out["status"]["error"] == v["error"]


def test_emit_kats_and_oracle(cro, oracle, kats):
    d = kats["emit_derived"]
    assert cro.emit_fm_scale_up(*d["fm_scale_up"]["args"]) == d["fm_scale_up"]["json"]
    assert cro.emit_fm_scale_down(*d["fm_scale_down"]["args"]) == d["fm_scale_down"]["json"]
    assert cro.emit_cm_scale_up(*d["cm_scale_up"]["args"]) == d["cm_scale_up"]["json"]
    assert cro.emit_cm_scale_down(*d["cm_scale_down"]["args"]) == d["cm_scale_down"]["json"]
    assert cro.emit_sunfish_request(*d["sunfish"]["args"]) == d["sunfish"]["json"]
    assert cro.emit_status_json(*d["status_online"]["args"]) == d["status_online"]["json"]
    for v in kats["normalize"]:
        assert cro.normalize(v["kind"], v["in"]) == v["out"], v["cite"]


def test_emit_fuzz_vs_oracle(cro, oracle):
    rng = random.Random(5)
    for _ in range(1500):
        a = [rand_text(rng, NASTY, rng.randrange(0, 12)).replace("\x00", "") for _ in range(5)]
        assert cro.emit_status_json(*a[:4]) == oracle.emit_status(*a[:4])
        assert cro.emit_scalar_status_json(*a) == oracle.emit_scalar_status(*a)
        assert cro.emit_fm_scale_up(*a[:4]) == oracle.emit_fm_scale_up(*a[:4])
        assert cro.emit_fm_scale_down(*a[:4]) == oracle.emit_fm_scale_down(*a[:4])
        n = rng.randrange(-3, 1000)
        assert cro.emit_cm_scale_up(a[0], n) == oracle.emit_cm_scale_up(a[0], n)
        assert cro.emit_cm_scale_down(a[0], n, a[1]) == oracle.emit_cm_scale_down(a[0], n, a[1])
        assert cro.emit_sunfish_request(a[0], n, a[1], a[2]) == oracle.emit_sunfish(a[0], n, a[1], a[2])
    # invalid UTF-8 goes out as one � per bad byte, like encoding/json
    for raw in (b"\xff", b"a\xc3", b"\xe2\x82", b"\xed\xa0\x80", b"\xc0\xaf", b"\xf4\x90\x80\x80"):
        s = raw.decode("utf-8", "surrogateescape")
        assert cro.emit_status_json(s) == oracle.emit_status(s)


def test_parse_fuzz_vs_oracle(cro, oracle):
    rng = random.Random(11)
    queries = ["gpu_uuid", "device_minor,gpu_uuid,pci.bus_id", "a,b", "gpu_uuid,gpu_uuid", " gpu_uuid , pci.bus_id"]
    for _ in range(3000):
        so = rand_text(rng, NASTY, rng.randrange(0, 40))
        se = "" if rng.random() < 0.8 else rand_text(rng, NASTY, rng.randrange(1, 8))
        ee = None if rng.random() < 0.85 else "exit status 9"
        if rng.random() < 0.05:
            so = " No devices were found\n"
        q = rng.choice(queries)
        r = oracle.parse_gpu_csv(so, se, ee, q)
        rc, text = cro.getGPUInfoFromNvidiaSmiOutput(so, se, ee, q)
        assert (rc, text) == (r.code, r.to_json() if r.code == 0 else r.error), (so, se, ee, q)
        r = oracle.parse_proc_csv(so, se, ee, q)
        rc, text = cro.getGPUInfoFromProcOutput(so, se, ee, q)
        assert (rc, text) == (r.code, r.to_json() if r.code == 0 else r.error), (so, se, ee, q)


def test_multi_gpu_csv_round_trip(cro, oracle):
    """emit_csv -> the reference's parse rule -> the same identities (8-GPU box, ragged: 0 devices)."""
    devs = []
    for i in range(8):
        d = cro.DevInfo()
        d.cuda_ordinal, d.device_minor = i, 7 - i
        d.gpu_uuid = ("GPU-%08x-0000-1111-2222-333333333333" % i).encode()
        d.pci_bus_id = ("00000000:%02X:00.0" % (0x1B + 16 * i)).encode()
        d.name = b"NVIDIA B200"
        devs.append(d)
    text = cro.emit_csv(devs, "device_minor,gpu_uuid,pci.bus_id")
    assert text.splitlines()[0] == "7, GPU-00000000-0000-1111-2222-333333333333, 00000000:1B:00.0"
    r = oracle.parse_gpu_csv(text, "", None, "device_minor,gpu_uuid,pci.bus_id")
    assert [m["gpu_uuid"] for m in r.infos] == [d.gpu_uuid.decode() for d in devs]
    assert [m["device_minor"] for m in r.infos] == [str(7 - i) for i in range(8)]
    rc, js = cro.getGPUInfoFromNvidiaSmiOutput(text, "", None, "device_minor,gpu_uuid,pci.bus_id")
    assert rc == 0 and js == r.to_json()
    assert cro.emit_csv([], "gpu_uuid") == "No devices were found\n"
    assert cro.getGPUInfoFromNvidiaSmiOutput("No devices were found\n", "ignored", "ignored", "gpu_uuid") == (0, "[]")
    assert cro.CheckGPUVisible(devs, devs[3].gpu_uuid.decode())
    assert not cro.CheckGPUVisible(devs, "GPU-not-there")
    assert not cro.CheckGPUVisible([], "GPU-not-there")
    with pytest.raises(cro.ProbeError):
        cro.emit_csv(devs, "gpu_uuid,bogus_field")


def test_attach_fuzz_vs_oracle(cro, oracle):
    import __graft_entry__ as g
    rng = random.Random(1234)
    ds = ["nvidia-gpu-operator/nvidia-device-plugin-daemonset", "nvidia-gpu-operator/nvidia-dcgm",
          "nvidia-dra-driver-gpu/nvidia-dra-driver-gpu-kubelet-plugin"]
    for _ in range(1500):
        inp = oracle.AttachInput(
            deleting=rng.random() < 0.2, device_resource_type=rng.choice(["DEVICE_PLUGIN", "DRA"]),
            provider_waiting=rng.random() < 0.1,
            provider_error=rng.choice(["", "", "", "boom <x> & \"y\"", "runtime error: slice bounds out of range [:1] with length 0"]),
            provider_device_id=rng.choice(["GPU-aaaa", "GPU-bbbb"]), provider_cdi_device_id="res-1",
            std_out=rng.choice(["", "GPU-aaaa\nGPU-bbbb\n", " GPU-cccc ", "No devices were found\n", "GPU-aaaa\n\nGPU-cccc"]),
            std_err=rng.choice(["", "", "", "oops"]), exec_err=rng.choice([None, None, None, "exit status 1"]),
            driver_pod_missing=rng.random() < 0.1,
            ds_err=rng.choice([{}, {}, {ds[1]: "daemonsets.apps \"nvidia-dcgm\" not found"}, {ds[2]: "x"}, {ds[0]: "y", ds[1]: "z"},
                               {ds[1]: "runtime error: invalid memory address or nil pointer dereference"}]),
            slice_uuids=rng.choice([None, None, [], ["GPU-aaaa"], ["GPU-bbbb", "GPU-cccc"]]),
            update_fail_after=rng.choice([None, None, None, 0, 1, 2]), update_fail_error="Operation cannot be fulfilled on composableresources")
        st = oracle.Status("Attaching", rng.choice(["", "old error"]), rng.choice(["", "GPU-aaaa", "GPU-zzzz"]), rng.choice(["", "res-0"]))
        want_st, want_rq, want_err, want_n = oracle.attach_step(inp, st)
        req = {"name": inp.name, "spec": {"type": "gpu", "model": "m", "target_node": inp.target_node},
               "status": {"state": st.state, "error": st.error, "device_id": st.device_id, "cdi_device_id": st.cdi_device_id},
               "deleting": inp.deleting, "device_resource_type": inp.device_resource_type, "probe": False,
               "provider": {"waiting": inp.provider_waiting, "error": inp.provider_error,
                            "device_id": inp.provider_device_id, "cdi_device_id": inp.provider_cdi_device_id},
               "enumeration": {"stdout": inp.std_out, "stderr": inp.std_err, "exec_err": inp.exec_err},
               "driver_pod_missing": inp.driver_pod_missing, "daemonset_errors": inp.ds_err}
        if inp.slice_uuids is not None:
            req["resource_slices"] = [{"devices": [{"attributes": {"uuid": u}} for u in inp.slice_uuids]}]
        if inp.update_fail_after is not None:
            req["status_update_failures"] = {"after": inp.update_fail_after, "error": inp.update_fail_error}
        out = cro.reconcile_attach(None, req)
        assert g.json_status(out) == want_st.to_json(), (inp, st, out["_raw"])
        assert out["requeue_after_s"] == want_rq and out["error"] == want_err, (inp, st, out["_raw"])
        assert len(out["status_updates"]) == want_n, (inp, st, out["_raw"])


def test_fm_response_parse(cro, oracle, kats):
    fx = kats["fixtures"]
    for key, body in fx["fm_update_body"].items():
        if key.startswith("_"):
            continue
        assert cro.fm_parse_scale_up_response(body, fx["name"], "gpu", fx["model"]) == \
            oracle.fm_scale_up_response_to_ids(body, fx["name"], "gpu", fx["model"]), key
    # wrong model / wrong type fall through to "can not find"
    assert cro.fm_parse_scale_up_response(fx["fm_update_body"]["isAdded"], "n", "gpu", "other")[2] == \
        "can not find the added gpu when using FM to add gpu"
    assert cro.fm_parse_scale_up_response(fx["fm_update_body"]["isAdded"], "n", "cxlmemory", fx["model"])[2] == \
        "can not find the added gpu when using FM to add gpu"


def test_probe_annotations_are_additive(cro):
    r = cro.ProbeResult()
    r.status, r.sweep_bytes, r.read_best_ns, r.fill_ns = 0, 4 << 30, 600000, 700000
    r.gpu_uuid = b"GPU-x"
    r.checksum_xor, r.checksum_sum, r.checksum_wsum = 0x1234, 0xabcd, 0x77
    r.nonce, r.copy_sweeps, r.copy_verified = 3, 5, 5
    js = json.loads(cro.emit_probe_annotations_json(r))
    assert all(k.startswith("cohdi.io/probe-") for k in js)
    assert js["cohdi.io/probe-hbm-read-gbs"] == "7158.2"          # 4 GiB / 600 us, integer arithmetic
    assert js["cohdi.io/probe-checksum"] == "0000000000001234:000000000000abcd:0000000000000077"
    assert js["cohdi.io/probe-nonce"] == "3" and js["cohdi.io/probe-copies-verified"] == "5/5"
    assert list(js) == sorted(js)                                  # Go marshals map keys sorted


def test_go_panics_unwind_without_a_status_write(cro, oracle):
    """a Go panic is not an error.  parts[i] on a short CSV row (gpus.go:912-914) panics inside
    CheckGPUVisible: the reference unwinds past requeueOnErr (no Status().Update) and controller-runtime reports
    "panic: ... [recovered]" — so Status.Error must stay what it was."""
    base = {"name": "cr", "spec": {"type": "gpu", "model": "m", "target_node": "worker-0"}, "probe": False,
            "status": {"state": "Attaching", "error": "older", "device_id": "GPU-aaaa", "cdi_device_id": "res-0"},
            "device_resource_type": "DEVICE_PLUGIN"}
    # the FM reply's res_op_status is "": OptionStatus[:1] panics inside AddResource (fti/fm/client.go:195)
    out = cro.reconcile_attach(None, dict(base, status={"state": "Attaching"},
                                          provider={"error": "runtime error: slice bounds out of range [:1] with length 0"}))
    assert out["error"] == "panic: runtime error: slice bounds out of range [:1] with length 0 [recovered]"
    assert out["status_updates"] == [] and out["status"] == {"state": "Attaching"}
    # an ORDINARY provider error is written into Status.Error by requeueOnErr (:423-433)
    out = cro.reconcile_attach(None, dict(base, status={"state": "Attaching"}, provider={"error": "fabric said no"}))
    assert out["error"] == "fabric said no" and out["status"]["error"] == "fabric said no" and len(out["status_updates"]) == 1
    # Online: CheckResource's error is recorded, never returned — but a panic inside it is not an error
    on = dict(base, status={"state": "Online", "device_id": "GPU-aaaa", "cdi_device_id": "res-0"})
    out = cro.reconcile_attach(None, dict(on, provider={"check_resource_error": "runtime error: index out of range [0] with length 0"}))
    assert out["error"] == "panic: runtime error: index out of range [0] with length 0 [recovered]" and out["status_updates"] == []
    out = cro.reconcile_attach(None, dict(on, provider={"check_resource_error": "device is in Critical state"}))
    assert out["error"] == "" and out["status"]["error"] == "device is in Critical state" and out["requeue_after_s"] == 30


def test_a_refused_status_write_stops_the_handler_where_the_reference_stops(cro, oracle):
    """composableresource_controller.go:233-235: if the IDs cannot be stored, nothing after it runs — no daemonset
    restart, no visibility check — and requeueOnErr's own write of the error may fail too (only logged)."""
    req = {"name": "cr", "spec": {"type": "gpu", "model": "m", "target_node": "worker-0"}, "probe": False,
           "status": {"state": "Attaching"}, "device_resource_type": "DEVICE_PLUGIN",
           "provider": {"device_id": "GPU-aaaa", "cdi_device_id": "res-0"},
           "enumeration": {"stdout": "GPU-aaaa\n", "stderr": "", "exec_err": None},
           "daemonset_errors": {"nvidia-gpu-operator/nvidia-dcgm": "daemonsets.apps \"nvidia-dcgm\" not found"}}
    ok = cro.reconcile_attach(None, req)
    assert ok["status"]["state"] == "Online" and len(ok["status_updates"]) == 3 and ok["failed_status_updates"] == 0
    conflict = "Operation cannot be fulfilled on composableresources.cro.hpsys.ibm.ie.com \"cr\": the object has been modified"
    out = cro.reconcile_attach(None, dict(req, status_update_failures={"after": 0, "error": conflict}))
    assert out["error"] == conflict and out["status"]["state"] == "Attaching"
    assert len(out["status_updates"]) == 2 and out["failed_status_updates"] == 2       # the IDs, then requeueOnErr's own attempt
    assert "daemonset_restarts" not in out or out["daemonset_restarts"] == []           # never got there
    out = cro.reconcile_attach(None, dict(req, status_update_failures={"after": 2, "error": conflict}))
    assert out["error"] == conflict and out["status"]["state"] == "Online"              # in memory; the write of "Online" was refused
    assert len(out["status_updates"]) == 3 and out["failed_status_updates"] == 1
