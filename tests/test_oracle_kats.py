"""Pins the oracle before anything trusts it: golden vectors from the reference's
own tests (tests/golden/reference_kats.json), committed pattern vectors, and
C-oracle vs Python-oracle agreement on seeded adversarial inputs."""
import json
import os
import random

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rand_text(rng, alphabet, n):
    return "".join(rng.choice(alphabet) for _ in range(n))


NASTY = ["a", "B", "0", "9", ",", ",", " ", "\t", "\n", "\n", "\r", "\v", "\f", "<", ">", "&", '"', "\\",
         "\u00a0", "\u0085", "\u2028", "\u2029", "\u3000", "\u2003", "\u00e9", "\u65e5", "\U0001f600", "\x01", "\x7f", "-", ":", ".", "G", "P", "U"]


def test_parse_kats_both_oracles(oracle, coracle, kats):
    for v in kats["parse"]:
        r = oracle.parse_gpu_csv(v["stdout"], v["stderr"], v["exec_err"], v["query"])
        rc, text = coracle.parse_gpu_csv(v["stdout"], v["stderr"], v["exec_err"], v["query"])
        assert r.code == v["code"] == rc, v["cite"]
        if v["code"] == 0:
            assert r.to_json() == v["json"] == text, v["cite"]
        else:
            assert r.error == v["error"] == text, v["cite"]


def _attach_inputs(oracle, kats, v):
    fx = kats["fixtures"]
    inp = oracle.AttachInput(name=fx["name"], target_node=fx["node"], deleting=v.get("deleting", False),
                             device_resource_type=v["device_resource_type"])
    if "fm_update" in v:
        dev, cdi, err = oracle.fm_scale_up_response_to_ids(fx["fm_update_body"][v["fm_update"]], fx["name"], "gpu", fx["model"])
        inp.provider_device_id, inp.provider_cdi_device_id, inp.provider_error = dev, cdi, err
    else:
        p = v.get("provider", {})
        inp.provider_waiting = p.get("waiting", False)
        inp.provider_error = p.get("error", "")
        inp.provider_device_id, inp.provider_cdi_device_id = p.get("device_id", ""), p.get("cdi_device_id", "")
    en = v.get("enumeration", {})
    inp.std_out, inp.std_err, inp.exec_err = en.get("stdout", ""), en.get("stderr", ""), en.get("exec_err")
    inp.driver_pod_missing = v.get("driver_pod_missing", False)
    inp.ds_err = dict(v.get("daemonset_errors", {}))
    if "resource_slices" in v:
        inp.slice_uuids = [d["attributes"]["uuid"] for rs in v["resource_slices"] for d in rs.get("devices", [])
                           if "uuid" in d.get("attributes", {})]
    si = v["status_in"]
    st = oracle.Status(si.get("state", ""), si.get("error", ""), si.get("device_id", ""), si.get("cdi_device_id", ""))
    return inp, st


def test_attach_kats_both_oracles(oracle, coracle, kats):
    for v in kats["attach"]:
        inp, st = _attach_inputs(oracle, kats, v)
        for impl in (oracle.attach_step, coracle.attach_step):
            out, rq, err, _n = impl(inp, st)
            exp = v["expected"]
            assert err == exp.get("error", ""), (v["cite"], err)
            if "status" in exp:
                e = exp["status"]
                want = oracle.Status(e.get("state", ""), e.get("error", ""), e.get("device_id", ""), e.get("cdi_device_id", ""))
                assert out == want, (v["cite"], out)
                assert rq == exp["requeue_after_s"], v["cite"]


def test_fm_gate_kats(oracle, coracle, kats):
    fx = kats["fixtures"]
    body = fx["fm_update_body"]
    assert oracle.fm_scale_up_response_to_ids(body["isAdded"], fx["name"], "gpu", fx["model"]) == (fx["device_id"], fx["cdi_device_id"], "")
    assert oracle.fm_scale_up_response_to_ids(body["isAddedWarning"], fx["name"], "gpu", fx["model"])[2] == ""
    assert coracle.fm_gate(fx["name"], "0") == (0, "")
    assert coracle.fm_gate(fx["name"], "1") == (0, "")
    assert coracle.fm_gate(fx["name"], "2") == (1, "the FM attached device called by test-composable-resource is in Critical state in FM")
    assert coracle.fm_gate(fx["name"], "3") == (1, "the FM attached device called by test-composable-resource is in unknown state '3' in FM")


def test_normalize_and_emit_kats(oracle, coracle, kats):
    for v in kats["normalize"]:
        assert oracle.normalize(v["kind"], v["in"]) == v["out"] == coracle.normalize(v["kind"], v["in"]), v["cite"]
    d = kats["emit_derived"]
    assert oracle.emit_fm_scale_up(*d["fm_scale_up"]["args"]) == d["fm_scale_up"]["json"] == coracle.emit_fm_scale_up(*d["fm_scale_up"]["args"])
    assert oracle.emit_fm_scale_down(*d["fm_scale_down"]["args"]) == d["fm_scale_down"]["json"] == coracle.emit_fm_scale_down(*d["fm_scale_down"]["args"])
    assert oracle.emit_cm_scale_up(*d["cm_scale_up"]["args"]) == d["cm_scale_up"]["json"] == coracle.emit_cm_scale_up(*d["cm_scale_up"]["args"])
    assert oracle.emit_cm_scale_down(*d["cm_scale_down"]["args"]) == d["cm_scale_down"]["json"] == coracle.emit_cm_scale_down(*d["cm_scale_down"]["args"])
    assert oracle.emit_sunfish(*d["sunfish"]["args"]) == d["sunfish"]["json"] == coracle.emit_sunfish(*d["sunfish"]["args"])
    assert oracle.emit_status(*d["status_online"]["args"]) == d["status_online"]["json"] == coracle.emit_status(*d["status_online"]["args"])
    # every emitted document must be valid JSON
    for k, v in d.items():
        if isinstance(v, dict):
            json.loads(v["json"])


def test_pattern_kats(oracle, coracle):
    with open(os.path.join(ROOT, "tests", "golden", "pattern_kats.json")) as f:
        g = json.load(f)
    for w in g["words"]:
        seed, i, word = int(w["seed"], 16), int(w["i"]), int(w["word"], 16)
        assert oracle.pattern_word(seed, i) == word == coracle.pattern_word(seed, i)
    for c in g["checksums"]:
        seed = int(c["seed"], 16)
        want = (int(c["xor"], 16), int(c["sum"], 16), int(c["wsum"], 16))
        assert coracle.checksum(seed, c["first"], c["n_words"]) == want
        assert oracle.checksum_np(seed, c["first"], c["n_words"]) == want
    for p in g["probe_seeds"]:
        want = int(p["seed"], 16)
        assert oracle.probe_seed(int(p["seed_base"], 16), p["minor"], p["nonce"]) == want
        assert coracle.probe_seed(int(p["seed_base"], 16), p["minor"], p["nonce"]) == want
    # std::mt19937_64's check value from the C++ standard ([rand.predef]): the generator restated in the KAT script is right
    assert g["mt19937_64_10000th_of_default_seed"] == "9981545732273789042"
    for c in g["chase_ends"]:
        assert coracle.chase_end(c["minor_src"], c["minor_dst"], c["hops"]) == c["end"]
        if c["hops"] == 65536:
            assert c["end"] == 0          # a Sattolo permutation is ONE cycle through all 65536 slots


def test_checksum_properties(coracle):
    seed = 0x00C0FFEE00000003
    n = 1 << 18
    whole = coracle.checksum(seed, 0, n)
    # additivity over a split (XOR / wrapping add are associative)
    a, b = coracle.checksum(seed, 0, 12345), coracle.checksum(seed, 12345, n - 12345)
    M = (1 << 64) - 1
    assert (a[0] ^ b[0], (a[1] + b[1]) & M, (a[2] + b[2]) & M) == whole
    # threads do not change the result
    assert coracle.checksum(seed, 0, n, threads=7) == whole
    # empty range
    assert coracle.checksum(seed, 0, 0) == (0, 0, 0)
    # the position-weighted component sees what XOR and sum cannot: two words swapping places
    import ctypes
    buf = (ctypes.c_uint64 * 1024)()
    coracle.lib.oracle_fill(buf, seed, 0, 1024)
    def fold(b):
        x, s, w = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        coracle.lib.oracle_checksum_buffer(b, 1024, ctypes.byref(x), ctypes.byref(s), ctypes.byref(w))
        return x.value, s.value, w.value
    clean = fold(buf)
    assert clean == coracle.checksum(seed, 0, 1024)
    buf[17], buf[900] = buf[900], buf[17]
    swapped = fold(buf)
    assert swapped[:2] == clean[:2] and swapped[2] != clean[2]


def test_c_vs_python_parse_fuzz(oracle, coracle):
    rng = random.Random(20260921)
    queries = ["gpu_uuid", "device_minor,gpu_uuid,pci.bus_id", "a,b", "gpu_uuid,gpu_uuid", " gpu_uuid , pci.bus_id", "device_minor"]
    for _ in range(3000):
        so = rand_text(rng, NASTY, rng.randrange(0, 40))
        se = "" if rng.random() < 0.8 else rand_text(rng, NASTY, rng.randrange(1, 8))
        ee = None if rng.random() < 0.85 else rand_text(rng, "abc <>", 5)
        if rng.random() < 0.05:
            so = rng.choice([" \n", "\t", ""]) + "No devices were found" + rng.choice(["", "\n", " \r\n"])
        q = rng.choice(queries)
        for py, c in ((oracle.parse_gpu_csv, coracle.parse_gpu_csv), (oracle.parse_proc_csv, coracle.parse_proc_csv)):
            r = py(so, se, ee, q)
            rc, text = c(so, se, ee, q)
            assert rc == r.code, (so, se, ee, q, text, r)
            assert text == (r.to_json() if r.code == 0 else r.error), (so, se, ee, q)


def test_c_vs_python_json_string_fuzz(oracle, coracle):
    rng = random.Random(7)
    for _ in range(3000):
        s = rand_text(rng, NASTY, rng.randrange(0, 24))
        assert coracle.json_string(s) == oracle.go_json_string(s), repr(s)
    # invalid UTF-8: one � per offending byte (utf8.DecodeRune returns width 1)
    for raw in (b"\xff", b"a\xc3", b"\xe2\x82", b"\xed\xa0\x80", b"\xc0\xaf", b"\xf4\x90\x80\x80", b"ok\xf0\x9f\x98\x80\x80"):
        s = raw.decode("utf-8", "surrogateescape")
        assert coracle.json_string(s) == oracle.go_json_string(s), raw


def test_c_vs_python_attach_fuzz(oracle, coracle):
    rng = random.Random(99)
    uuids = ["GPU-aaaa", "GPU-bbbb", "GPU-cccc"]
    for _ in range(2000):
        inp = oracle.AttachInput(
            deleting=rng.random() < 0.2,
            device_resource_type=rng.choice(["DEVICE_PLUGIN", "DRA"]),
            provider_waiting=rng.random() < 0.1,
            provider_error=rng.choice(["", "", "", "boom <x>", "runtime error: slice bounds out of range [:1] with length 0"]),
            provider_device_id=rng.choice(uuids), provider_cdi_device_id="res-1",
            std_out=rng.choice(["", "GPU-aaaa\nGPU-bbbb\n", " GPU-cccc ", "No devices were found\n", "GPU-aaaa\n\nGPU-cccc"]),
            std_err=rng.choice(["", "", "", "oops"]),
            exec_err=rng.choice([None, None, None, "exit status 1"]),
            driver_pod_missing=rng.random() < 0.1,
            ds_err=rng.choice([{}, {}, {"nvidia-gpu-operator/nvidia-dcgm": "daemonsets.apps \"nvidia-dcgm\" not found"},
                               {"nvidia-dra-driver-gpu/nvidia-dra-driver-gpu-kubelet-plugin": "x"},
                               {"nvidia-gpu-operator/nvidia-dcgm": "runtime error: invalid memory address or nil pointer dereference"}]),
            slice_uuids=rng.choice([None, None, [], ["GPU-aaaa"], ["GPU-bbbb", "GPU-cccc"]]),
            update_fail_after=rng.choice([None, None, None, 0, 1, 2]), update_fail_error="Operation cannot be fulfilled")
        st = oracle.Status("Attaching", rng.choice(["", "old error"]), rng.choice(["", "GPU-aaaa", "GPU-zzzz"]), rng.choice(["", "res-0"]))
        assert oracle.attach_step(inp, st) == coracle.attach_step(inp, st), (inp, st)


def test_proc_information(oracle, coracle):
    info = ("Model: \t\t NVIDIA B200\nIRQ:   \t\t 89\nGPU UUID: \t GPU-5e4ad8a4-c9d2-9d2a-3e0f-0a1b2c3d4e5f\n"
            "Video BIOS: \t 97.00.5e.00.01\nBus Type: \t PCIe\nDMA Size: \t 52 bits\nDMA Mask: \t 0xfffffffffffff\n"
            "Bus Location: \t 0000:1b:00.0\nDevice Minor: \t 3\nGPU Excluded:\t No\n")
    want = "3,GPU-5e4ad8a4-c9d2-9d2a-3e0f-0a1b2c3d4e5f,0000:1b:00.0\n"
    assert oracle.proc_information_to_line(info) == want == coracle.proc_information_to_line(info)
    broken = info.replace("Device Minor: \t 3\n", "Device Minor:\n")
    assert oracle.proc_information_to_line(broken) == "" == coracle.proc_information_to_line(broken)
    r = oracle.parse_proc_csv(want, "", None, "device_minor,gpu_uuid,pci.bus_id")
    assert r.to_json() == '[{"device_minor":"3","gpu_uuid":"GPU-5e4ad8a4-c9d2-9d2a-3e0f-0a1b2c3d4e5f","pci.bus_id":"0000:1b:00.0"}]'
    assert oracle.normalize(0, "0000:1b:00.0") == "0000:1B:00.0"
    assert oracle.parse_proc_csv("1,GPU-x", "", None, "gpu_uuid").error == "unexpected GPU information format: '1,GPU-x'"
    assert oracle.parse_proc_csv(want, "", None, "name").error == "unsupported field 'name' requested in queryArgs"
