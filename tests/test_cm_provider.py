"""CM flavour of the ID production (checkAddingResources, internal/cdi/fti/cm/client.go:432-509) against
the reference's generateCMMachineData fixture shapes (composableresource_controller_test.go:112-366)."""
import json
import random

DEV, RES = "GPU-device00-uuid-temp-0000-000000000000", "GPU-device00-uuid-temp-0000-000000000res"
MODEL = "NVIDIA-A100-PCIE-80GB"


def cm_machine_data(devices):
    """json.Marshal of fticmapi.MachineData as generateCMMachineData builds it (nil slices -> null)."""
    def detail(res, op):
        return {"fabric_uuid": "", "fabric_id": 0, "res_uuid": res, "fabr_gid": "", "res_type": "", "res_name": "",
                "res_status": "", "res_op_status": op,
                "resspecs": [{"resspec_uuid": "", "productname": "", "model": "", "vendor": "", "removable": True}],
                "tenant_uuid": "", "mach_uuid": ""}

    def spec(uuid, typ, model, devs):
        conds = None if model is None else [{"column": "model", "operator": "eq", "value": model}]
        return {"spec_uuid": uuid, "type": typ, "selector": {"version": "", "expression": {"conditions": conds}},
                "min_resspec_count": 0, "max_resspec_count": 0, "device_count": 0, "devices": devs}
    devs = None if devices is None else [
        {"device_id": d[0], "status": d[1], "status_reason": d[2], "detail": detail(d[3], d[4])} for d in devices]
    specs = [spec("spec0000-uuid-temp-0000-000000000000", "cpu", None, None),
             spec("spec0000-uuid-temp-0000-000000000001", "gpu", "NVIDIA-ANOTHER-GPU", None),
             spec("spec0000-uuid-temp-0000-000000000002", "gpu", MODEL, devs)]
    return json.dumps({"data": {"tenant_uuid": "", "cluster": {"cluster_uuid": "", "machine": {
        "uuid": "", "name": "", "status": "", "status_reason": "", "resspecs": specs}}}}, separators=(",", ":"))


def test_cm_kats(cro, oracle):
    # :2064 "should wait when the GPU has not yet been added in CM": no device -> resize request for spec ...0002
    body = cm_machine_data(None)
    want = ("spec0000-uuid-temp-0000-000000000002", 0, "", "", "")
    assert cro.cm_check_adding_resources(body, [], "gpu", MODEL) == want == oracle.cm_check_adding_resources(body, [], "gpu", MODEL)
    assert cro.emit_cm_scale_up(want[0], want[1] + 1) == '{"increase_resource_count":{"spec_uuid":"spec0000-uuid-temp-0000-000000000002","device_count":1}}'
    # isSucceeded fixture: ADD_COMPLETE device no CR owns -> ids
    body = cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")])
    want = ("", 0, DEV, RES, "")
    assert cro.cm_check_adding_resources(body, [""], "gpu", MODEL) == want == oracle.cm_check_adding_resources(body, [""], "gpu", MODEL)
    # ... unless a ComposableResource already owns it -> back to the resize request
    assert cro.cm_check_adding_resources(body, [DEV], "gpu", MODEL)[0] == "spec0000-uuid-temp-0000-000000000002"
    # :2196 isAttachFailed fixture
    body = cm_machine_data([("GPU-device00-uuid-temp-fail-000000000000", "ADD_FAILED", "add failed due to some reasons",
                             "GPU-device00-uuid-temp-fail-000000000res", "0")])
    got = cro.cm_check_adding_resources(body, [], "gpu", MODEL)
    assert got == oracle.cm_check_adding_resources(body, [], "gpu", MODEL)
    assert got[4] == "an error occurred with the resource in CM: 'add failed due to some reasons'"
    # no matching spec at all
    assert cro.cm_check_adding_resources(body, [], "gpu", "NVIDIA-H100") == ("", 0, "", "", "")
    assert cro.cm_check_adding_resources(body, [], "cxlmemory", MODEL) == ("", 0, "", "", "")


def test_cm_through_the_attach_step(cro, oracle, kats):
    import __graft_entry__ as g
    base = {"name": "test-composable-resource", "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"},
            "status": {"state": "Attaching"}, "device_resource_type": "DRA", "probe": False,
            "enumeration": {"stdout": DEV, "stderr": ""},
            "resource_slices": [{"devices": [{"attributes": {"uuid": DEV}}]}]}
    out = cro.reconcile_attach(None, dict(base, provider={"cm_machine_body": cm_machine_data(None)}))
    assert g.json_status(out) == '{"state":"Attaching"}' and out["requeue_after_s"] == 30 and out["error"] == ""
    out = cro.reconcile_attach(None, dict(base, provider={"cm_machine_body": cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")])}))
    assert g.json_status(out) == oracle.emit_status("Online", "", DEV, RES)
    out = cro.reconcile_attach(None, dict(base, provider={"cm_machine_body": cm_machine_data(
        [("GPU-device00-uuid-temp-fail-000000000000", "ADD_FAILED", "add failed due to some reasons", "r", "0")])}))
    assert out["error"] == "an error occurred with the resource in CM: 'add failed due to some reasons'"


def test_cm_fuzz_vs_oracle(cro, oracle):
    rng = random.Random(3)
    ids = ["GPU-a", "GPU-b", "GPU-c", "GPU-d"]
    for _ in range(800):
        devs = None if rng.random() < 0.15 else [
            (rng.choice(ids), rng.choice(["ADD_COMPLETE", "ADD_FAILED", "ADDING", "REMOVE_FAILED", ""]),
             rng.choice(["", "why <not>"]), "res-" + rng.choice("xyz"), rng.choice(["0", "1", "2"]))
            for _ in range(rng.randrange(0, 4))]
        body = cm_machine_data(devs)
        existing = rng.sample(ids, rng.randrange(0, 4))
        model = rng.choice([MODEL, MODEL, "NVIDIA-ANOTHER-GPU", "nope"])
        typ = rng.choice(["gpu", "gpu", "cpu", "cxlmemory"])
        assert cro.cm_check_adding_resources(body, existing, typ, model) == oracle.cm_check_adding_resources(body, existing, typ, model)
