"""Fabric wire codec, response side: CheckResource / GetResources decisions over the
GET-machine bodies.  KATs: the Online-state entries of composableresource_controller_test.go
(CM :3633,3710,3787,3864; FM :8088,8166,8243,8320) on the generateFMMachineData / generateCMMachineData shapes."""
import json
import random

from test_cm_provider import cm_machine_data

DEV, RES = "GPU-device00-uuid-temp-0000-000000000000", "GPU-device00-uuid-temp-0000-000000000res"
MODEL = "NVIDIA-A100-PCIE-80GB"


def fm_machine_data(extra):
    """json.Marshal of ftifmapi.GetMachineResponse as generateFMMachineData builds it (:368-490)."""
    def res(uuid, typ, op, serial, model):
        cond = None if model is None else [{"column": "model", "operator": "eq", "value": model}]
        return {"res_uuid": uuid, "res_name": "", "res_type": typ, "res_status": 0, "res_op_status": op,
                "res_serial_num": serial, "res_spec": {"condition": cond}}
    resources = [res("device00-uuid-temp-0000-other0000000", "memory", "0", "", None),
                 res("GPU-device00-uuid-temp-0000-other0000000", "gpu", "0", "", "NVIDIA-OTHER")]
    resources += [res(*e) for e in extra]
    return json.dumps({"data": {"machines": [{"fabric_uuid": "", "fabric_id": 0, "mach_uuid": "", "mach_id": 0, "mach_name": "",
                                              "tenant_uuid": "", "mach_status": 0, "mach_status_detail": "",
                                              "resources": resources}]}}, separators=(",", ":"))


FM_KATS = [  # (op status of the device, expected Status.Error)
    (None, "the target device '%s' cannot be found in CDI system" % DEV),          # :8088
    ("0", ""),
    ("1", "the target gpu '%s' is showing a Warning status in FM" % DEV),           # :8166
    ("2", "the target gpu '%s' is showing a Critical status in FM" % DEV),          # :8243
    ("3", "the target gpu '%s' has unknown status '3' in FM" % DEV),                # :8320
]
CM_KATS = [
    (None, "the target device '%s' cannot be found in CDI system" % DEV),          # :3633
    ("0", ""),
    ("1", "the target gpu '%s' is showing a Warning status in CM" % DEV),           # :3710
    ("2", "the target gpu '%s' is showing a Critical status in CM" % DEV),          # :3787
    ("3", "the target gpu '%s' has unknown status '3' in CM" % DEV),                # :3864
]


def test_fm_check_resource_kats(cro, oracle):
    for op, want in FM_KATS:
        body = fm_machine_data([] if op is None else [(RES, "gpu", op, DEV, MODEL)])
        assert cro.fabric_check_resource("fm", body, "gpu", MODEL, DEV) == want == oracle.fabric_check_resource("fm", body, "gpu", MODEL, DEV)


def test_cm_check_resource_kats(cro, oracle):
    for op, want in CM_KATS:
        body = cm_machine_data(None if op is None else [(DEV, "ADD_COMPLETE", "", RES, op)])
        assert cro.fabric_check_resource("cm", body, "gpu", MODEL, DEV) == want == oracle.fabric_check_resource("cm", body, "gpu", MODEL, DEV)


def test_get_resources_feed_the_syncer(cro, oracle):
    body = fm_machine_data([(RES, "gpu", "0", DEV, MODEL)])
    got = cro.fabric_get_resources("fm", body, "worker-0", "machine0-uuid-temp-0000-000000000000")
    assert got == oracle.fabric_get_resources("fm", body, "worker-0", "machine0-uuid-temp-0000-000000000000")
    assert got[-1] == {"node_name": "worker-0", "machine_uuid": "machine0-uuid-temp-0000-000000000000", "device_type": "gpu",
                       "model": MODEL, "device_id": DEV, "cdi_device_id": RES}
    assert [g["model"] for g in got] == ["NVIDIA-OTHER", MODEL]           # the memory resource is skipped
    cm = cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")])
    got_cm = cro.fabric_get_resources("cm", cm, "worker-0", "m")
    assert got_cm == oracle.fabric_get_resources("cm", cm, "worker-0", "m")
    assert got_cm == [{"node_name": "worker-0", "machine_uuid": "m", "device_type": "gpu", "model": "", "device_id": DEV, "cdi_device_id": RES}]
    # end to end: GetResources -> UpstreamSyncer tick -> detach CR after the grace period
    with cro.Cluster({"nodes": ["worker-0"], "uuids": [DEV]}) as c:
        c.sync_upstream(got_cm, 1000)
        c.sync_upstream(got_cm, 1000 + 11 * 60)
        crs = [r for r in c.dump()["resources"].values() if r["labels"].get("cohdi.io/ready-to-detach-device-id") == DEV]
        assert len(crs) == 1 and crs[0]["labels"]["cohdi.io/ready-to-detach-cdi-device-id"] == RES


def test_fabric_fuzz_vs_oracle(cro, oracle):
    rng = random.Random(31)
    ids = ["GPU-a", "GPU-b", DEV]
    for _ in range(600):
        extra = [(rng.choice(["r1", "r2"]), rng.choice(["gpu", "gpu", "memory"]), rng.choice(["0", "1", "2", "3", "10", "x"]),
                  rng.choice(ids), rng.choice([MODEL, MODEL, "NVIDIA-OTHER", None])) for _ in range(rng.randrange(0, 4))]
        body = fm_machine_data(extra)
        dev, model, typ = rng.choice(ids), rng.choice([MODEL, "NVIDIA-OTHER"]), rng.choice(["gpu", "memory"])
        assert cro.fabric_check_resource("fm", body, typ, model, dev) == oracle.fabric_check_resource("fm", body, typ, model, dev)
        assert cro.fabric_get_resources("fm", body, "n", "m") == oracle.fabric_get_resources("fm", body, "n", "m")
        devs = None if rng.random() < 0.2 else [(rng.choice(ids), "ADD_COMPLETE", "", "res", rng.choice(["0", "1", "2", "7"])) for _ in range(rng.randrange(0, 3))]
        cm = cm_machine_data(devs)
        assert cro.fabric_check_resource("cm", cm, typ, model, dev) == oracle.fabric_check_resource("cm", cm, typ, model, dev)
        assert cro.fabric_get_resources("cm", cm, "n", "m") == oracle.fabric_get_resources("cm", cm, "n", "m")
    # the reference's unguarded indexing, kept as errors rather than crashes
    assert cro.fabric_check_resource("fm", '{"data":{"machines":[]}}', "gpu", MODEL, DEV) == "runtime error: index out of range [0] with length 0"
    empty_op = fm_machine_data([(RES, "gpu", "", DEV, MODEL)])
    assert cro.fabric_check_resource("fm", empty_op, "gpu", MODEL, DEV) == "runtime error: slice bounds out of range [:1] with length 0"
