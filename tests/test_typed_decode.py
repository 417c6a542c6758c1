"""A fabric reply whose JSON types do not fit the reference's wire structs (fm/api/get.go, fm/api/scale_up.go,
cm/api/machine.go) must fail the way json.Unmarshal fails it — "failed to unmarshal ... : json: cannot unmarshal string
into Go struct field GetMachineItem.data.machines.fabric_id of type int" — not be read leniently.  C++
(gojson::decodesInto over csrc/gotypes.cpp, through the FM / CM clients) against the Python restatement
(oracle/fabric_clients.type_mismatch) on valid replies with one to three values swapped for another JSON type.
THAT it is an error follows from the reference's code (fm/client.go:185,506; cm/client.go:425); the wording is go1.24's
UnmarshalTypeError and has no reference vector."""
import copy
import importlib
import json
import os
import random
import sys

import pytest

from test_cm_provider import cm_machine_data
from test_fabric_codec import fm_machine_data

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import fabric_clients as fc  # noqa: E402

KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
DEV, RES, MODEL = "GPU-device00-uuid-temp-0000-000000000000", "GPU-device00-uuid-temp-0000-000000000res", "NVIDIA-A100-PCIE-80GB"
OBJECTS = {"nodes": {"worker-0": {"annotations": {"machine.openshift.io/machine": "ns/m0"}}},
           "metal3machines": {"ns/m0": {"annotations": {"metal3.io/BareMetalHost": "ns/b0"}}},
           "baremetalhosts": {"ns/b0": {"annotations": {"cluster-manager.cdi.io/machine": "mach"}}}}


@pytest.fixture(scope="module")
def cro():
    return importlib.import_module("composable-resource-operator_b200")


def run(cro, kind, state, body):
    fabric = {"objects": OBJECTS, "http": [{"path_contains": "", "status": 200, "body": body}]}
    online = state == "Online"
    out = cro.reconcile_attach(None, {
        "name": "cr", "spec": {"type": "gpu", "model": MODEL, "target_node": "worker-0"},
        "status": {"state": state, "device_id": DEV if online else "", "cdi_device_id": RES if online else ""},
        "deleting": False, "probe": False,
        "env": {"DEVICE_RESOURCE_TYPE": "DRA", "CDI_PROVIDER_TYPE": "FTI_CDI", "FTI_CDI_API_TYPE": kind, "FTI_CDI_TENANT_ID": "t", "FTI_CDI_CLUSTER_ID": "c"},
        "fabric": fabric, "enumeration": {"stdout": "", "stderr": ""}, "resource_slices": []})
    client = (fc.CMClient if kind == "CM" else fc.FMClient)(fc.Fabric(fabric), "t", "c")

    def unwrap(e):       # a Go panic in the client surfaces as controller-runtime's "panic: <text> [recovered]", with no status write
        if e.startswith("panic: ") and e.endswith(" [recovered]"):
            assert out["status_updates"] == [], out
            return e[len("panic: "):-len(" [recovered]")]
        return e
    if online:
        got = unwrap(out["error"]) if out["error"] else out["status"].get("error", "")
        return got, client.check("gpu", MODEL, "worker-0", DEV)
    d, c, err = client.add("cr", "gpu", MODEL, "worker-0")
    return unwrap(out["error"]), ("" if err == fc.ERR_ATTACHING else err)


def paths(v, at=()):
    """every value position of a JSON tree"""
    yield at
    if isinstance(v, dict):
        for k in v:
            yield from paths(v[k], at + (k,))
    elif isinstance(v, list):
        for i, x in enumerate(v):
            yield from paths(x, at + (i,))


def put(root, at, val):
    if not at:
        return val
    node = root
    for k in at[:-1]:
        node = node[k]
    node[at[-1]] = val
    return root


HAND = [
    ("FM", "Online", '{"data":{"machines":[{"fabric_id":"x"}]}}',
     "failed to unmarshal FM get machine response body into machineData: json: cannot unmarshal string into Go struct field GetMachineItem.data.machines.fabric_id of type int"),
    ("FM", "Online", '{"data":{"machines":{"a":1}}}',
     "failed to unmarshal FM get machine response body into machineData: json: cannot unmarshal object into Go struct field GetMachineData.data.machines of type []api.GetMachineItem"),
    ("FM", "Online", '{"data":{"machines":[{"resources":[{"res_spec":{"condition":[{"value":5}]}}]}]}}',
     "failed to unmarshal FM get machine response body into machineData: json: cannot unmarshal number into Go struct field ConditionItem.data.machines.resources.res_spec.condition.value of type string"),
    ("FM", "Attaching", '{"data":{"machines":[{"mach_id":1.5}]}}',
     "failed to unmarshal FM scaleup response body into scaleUpResponse. Original error: json: cannot unmarshal number 1.5 into Go struct field ScaleUpResponseMachineItem.data.machines.mach_id of type int"),
    ("FM", "Attaching", '{"DATA":[]}',
     "failed to unmarshal FM scaleup response body into scaleUpResponse. Original error: json: cannot unmarshal array into Go struct field ScaleUpResponse.data of type api.ScaleUpResponseData"),
    ("CM", "Online", '{"data":{"cluster":{"machine":{"resspecs":[{"devices":[{"detail":{"resspecs":[{"removable":"yes"}]}}]}]}}}}',
     "failed to unmarshal CM get machine response body into machineData: json: cannot unmarshal string into Go struct field DeviceResourceSpec.data.cluster.machine.resspecs.devices.detail.resspecs.removable of type bool"),
    ("CM", "Attaching", '{"data":{"cluster":{"machine":{"resspecs":[{"device_count":"2"}]}}}}',
     "failed to unmarshal CM get machine response body into machineData: json: cannot unmarshal string into Go struct field ResourceSpec.data.cluster.machine.resspecs.device_count of type int"),
    ("CM", "Online", '{"data":{"cluster":true}}',
     "failed to unmarshal CM get machine response body into machineData: json: cannot unmarshal bool into Go struct field Data.data.cluster of type api.Cluster"),
]


@pytest.mark.parametrize("i", range(len(HAND)))
def test_hand_cases(cro, i):
    kind, state, body, want = HAND[i]
    got, oracle = run(cro, kind, state, body)
    assert oracle == want
    assert got == want


def test_type_swaps(cro):
    rng = random.Random(99)
    seeds = {("CM", "Online"): [cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")])],
             ("CM", "Attaching"): [cm_machine_data(None), cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")])],
             ("FM", "Online"): [fm_machine_data([(RES, "gpu", "0", DEV, MODEL)])],
             ("FM", "Attaching"): [v for k, v in KATS["fixtures"]["fm_update_body"].items() if k != "_cite"]}
    swaps = ["s", 7, 1.5, True, None, [], [1], {}, {"k": "v"}, "0", -3, 10**30]
    tally = {"mismatch": 0, "clean": 0}
    for it in range(1600):
        (kind, state), bodies = rng.choice(sorted(seeds.items()))
        tree = json.loads(rng.choice(bodies))
        for _ in range(rng.choice([1, 1, 2, 3])):
            at = rng.choice(list(paths(tree)))
            cur = tree
            for k in at:
                cur = cur[k]
            if rng.random() < 0.5:                   # a value of the SAME JSON type, or null: still decodes
                val = (rng.choice(["", "other", None]) if isinstance(cur, str) else rng.choice([True, False]) if isinstance(cur, bool) else
                       rng.choice([0, 3, -1, None]) if isinstance(cur, int) else rng.choice([[], None]) if isinstance(cur, list) else
                       rng.choice([{}, None, {"unknown_key": [1, "x"]}]))
            else:
                val = rng.choice(swaps)
            tree = put(copy.deepcopy(tree), at, val)
        body = json.dumps(tree)
        if rng.random() < 0.2:                       # keys fold: the mismatch is still found, and named by the TAG
            body = body.replace('"machines"', '"MACHINES"').replace('"resspecs"', '"Resspecs"')
        got, oracle = run(cro, kind, state, body)
        assert got == oracle, (it, kind, state, body, got, oracle)
        tally["mismatch" if "cannot unmarshal" in oracle else "clean"] += 1
    assert tally["mismatch"] > 500 and tally["clean"] > 300, tally


# ---- repeated members: encoding/json merges, it does not "take the last one" -----------------------------------
def test_repeated_members_merge_like_go(cro):
    # two "data" objects: the second one's machines array is decoded OVER the first one's elements, so element 0 keeps the
    # resources it got first; the device is found although the last "data" alone does not hold it
    first = json.loads(fm_machine_data([(RES, "gpu", "0", DEV, MODEL)]))["data"]
    body = '{"data":%s,"DATA":{"machines":[{"mach_name":"renamed"}]}}' % json.dumps(first)
    got, oracle = run(cro, "FM", "Online", body)
    assert got == oracle == ""
    # ... and a shorter second array truncates: the machine is gone
    body = '{"data":%s,"data":{"machines":[]}}' % json.dumps(first)
    got, oracle = run(cro, "FM", "Online", body)
    assert got == oracle == "runtime error: index out of range [0] with length 0"
    # null leaves what was there
    body = '{"data":%s,"data":null}' % json.dumps(first)
    got, oracle = run(cro, "FM", "Online", body)
    assert got == oracle == ""
    # CM: the op status arrives in a second copy of the device list, merged into the first device
    cm = json.loads(cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")]))
    specs = cm["data"]["cluster"]["machine"]["resspecs"]
    at = [i for i, s in enumerate(specs) if s["devices"]][0]                 # the matching spec is not the first one:
    patch_specs = [None] * at + [{"devices": [{"detail": {"res_op_status": "2"}}]}] + [None] * (len(specs) - at - 1)
    patch = {"data": {"cluster": {"machine": {"resspecs": patch_specs}}}}     # null elements leave their slots alone
    body = json.dumps(cm)[:-1] + "," + json.dumps(patch)[1:]
    assert specs[at]["devices"][0]["detail"]["res_op_status"] == "0"
    got, oracle = run(cro, "CM", "Online", body)
    assert got == oracle == "the target gpu '%s' is showing a Critical status in CM" % DEV


def test_duplicated_member_fuzz(cro):
    """Valid replies where random subtrees are repeated under the same (or a case-variant) key, sometimes edited."""
    rng = random.Random(4242)
    seeds = {("CM", "Online"): cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")]),
             ("CM", "Attaching"): cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "0")]),
             ("FM", "Online"): fm_machine_data([(RES, "gpu", "0", DEV, MODEL)]),
             ("FM", "Attaching"): [v for k, v in KATS["fixtures"]["fm_update_body"].items() if k != "_cite"][-1]}

    def emit(v):
        """json.dumps that repeats some object members"""
        if isinstance(v, dict):
            parts = []
            for k, x in v.items():
                parts.append('"%s":%s' % (k, emit(x)))
                if rng.random() < 0.25:
                    y = copy.deepcopy(x)
                    r = rng.random()
                    if r < 0.3:
                        y = None
                    elif r < 0.6 and isinstance(y, list):
                        y = y[:rng.randrange(len(y) + 1)] + ([{}] if rng.random() < 0.3 else [])
                    elif r < 0.8 and isinstance(y, dict) and y:
                        y.pop(rng.choice(sorted(y)))
                    elif isinstance(y, str):
                        y = rng.choice([y, "", "2", "gpu", MODEL])
                    parts.append('"%s":%s' % (k.upper() if rng.random() < 0.3 else k, emit(y)))
            return "{" + ",".join(parts) + "}"
        if isinstance(v, list):
            return "[" + ",".join(emit(x) for x in v) + "]"
        return json.dumps(v)

    outcomes = set()
    for it in range(800):
        (kind, state), body = rng.choice(sorted(seeds.items()))
        text = emit(json.loads(body))
        json.loads(text)
        got, oracle = run(cro, kind, state, text)
        assert got == oracle, (it, kind, state, text, got, oracle)
        outcomes.add(oracle[:40])
    assert len(outcomes) >= 4, outcomes
