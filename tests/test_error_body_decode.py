"""Unmarshal(body, &api.ErrorBody{}) of a non-200 fabric reply (fm/api/common.go:31-40, cm/api/machine.go:95-103) the way
encoding/json walks it: members in input order, keys matched exactly or case-folded, duplicates overwrite, the FIRST type
mismatch is the error, json.RawMessage keeps the member's text.  C++ (through the FM / CM clients) against the Python
restatement on generated bodies.  The mismatch WORDING follows go1.24 and has no reference vector."""
import importlib
import json
import os
import random
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import fabric_clients as fc  # noqa: E402


@pytest.fixture(scope="module")
def cro():
    return importlib.import_module("composable-resource-operator_b200")


def attach(cro, kind, body, status=500):
    return cro.reconcile_attach(None, {
        "name": "cr", "spec": {"type": "gpu", "model": "m", "target_node": "worker-0"}, "status": {"state": "Attaching"},
        "deleting": False, "probe": False,
        "env": {"DEVICE_RESOURCE_TYPE": "DRA", "CDI_PROVIDER_TYPE": "FTI_CDI", "FTI_CDI_API_TYPE": kind, "FTI_CDI_TENANT_ID": "t", "FTI_CDI_CLUSTER_ID": "c"},
        "fabric": {"objects": {
            "nodes": {"worker-0": {"annotations": {"machine.openshift.io/machine": "ns/m0"}}},
            "metal3machines": {"ns/m0": {"annotations": {"metal3.io/BareMetalHost": "ns/b0"}}},
            "baremetalhosts": {"ns/b0": {"annotations": {"cluster-manager.cdi.io/machine": "mach"}}}},
            "http": [{"path_contains": "", "status": status, "body": body}]},
        "enumeration": {"stdout": "", "stderr": ""}, "resource_slices": []})["error"]


def want(kind, body):
    return fc.fm_error("scaleup", body) if kind == "FM" else fc.cm_error("get", body)


HAND = [
    ('{"status":500,"detail":{"code":"E1","message":"boom"}}', None),
    ('{"detail":{"code":5},"status":"x"}', "Go struct field ErrorDetail.detail.code of type string"),        # input order, not struct order
    ('{"status":"x","detail":{"code":5}}', "cannot unmarshal string into Go struct field ErrorBody.status of type int"),
    ('{"status":"x","status":7}', "cannot unmarshal string into Go struct field ErrorBody.status"),            # the error stays, 7 lands
    ('{"status":1.5}', "cannot unmarshal number 1.5 into Go struct field ErrorBody.status of type int"),
    ('{"detail":7}', "cannot unmarshal number into Go struct field ErrorBody.detail of type api.ErrorDetail"),
    ('{"detail":{"code":7}}', "cannot unmarshal number into Go struct field ErrorDetail.detail.code of type string"),
    ('{"detail":{"code":"A"},"DETAIL":{"message":"x"}}', None),                                              # two detail objects merge
    ('{"detail":{"Code":"A","code":null,"MESSAGE":"m"}}', None),
    ('{"detail":null,"status":null}', None),
    ('null', None),
]


@pytest.mark.parametrize("kind", ["FM", "CM"])
@pytest.mark.parametrize("i", range(len(HAND)))
def test_hand_cases(cro, kind, i):
    body, fragment = HAND[i]
    w = want(kind, body)
    assert attach(cro, kind, body) == w
    if fragment:
        assert fragment in w, w
    else:
        assert "Original error" not in w, w


def test_merged_details_and_raw_message(cro):
    assert attach(cro, "FM", '{"detail":{"code":"A"},"DETAIL":{"message":"x"}}') == \
        "failed to process FM scaleup request. FM returned code: 'A', error message: 'x'"
    assert attach(cro, "FM", '{"detail":{"message": {"k": [1, 2]} ,"message":null,"code":"Z"}}') == \
        "failed to process FM scaleup request. FM returned code: 'Z', error message: ''"
    assert attach(cro, "FM", '{"detail":{"message":"first","message": [ 1 ]}}') == \
        "failed to process FM scaleup request. FM returned code: '', error message: '[ 1 ]'"
    assert attach(cro, "CM", '{"status":3,"detail":{"message":"first","message":"last","code":"C"},"STATUS":4}') == \
        "failed to process CM get request. http returned status: '4', cm return code: 'C', error message: 'last'"


def test_generated_bodies(cro):
    rng = random.Random(7)
    scalars = ['"s"', "5", "-0", "1.5", "1e3", "99999999999999999999", "true", "null", "[]", "[1]", "{}", '{"a":1}', '"\\u00e9"', ' "pad" ']
    tally = {"ok": 0, "mismatch": 0}
    for it in range(1200):
        def detail():
            ms = []
            for _ in range(rng.randrange(0, 4)):
                k = rng.choice(["code", "Code", "message", "MESSAGE", "data", "Data", "other", "codſ"])
                ms.append('"%s":%s' % (k, rng.choice(scalars if rng.random() < 0.5 else ['"s"', '"t"', "null", "{}"])))
            return "{" + ",".join(ms) + "}"
        ms = []
        for _ in range(rng.randrange(0, 4)):
            k = rng.choice(["status", "Status", "detail", "DETAIL", "extra"])
            if k.lower() == "status":
                v = rng.choice(scalars if rng.random() < 0.4 else ["500", "404", "0"])
            elif k.lower() == "detail":
                v = detail() if rng.random() < 0.8 else rng.choice(scalars)
            else:
                v = rng.choice(scalars)
            ms.append('"%s":%s' % (k, v))
        body = "{" + ",".join(ms) + "}"
        json.loads(body)
        for kind in ("FM", "CM"):
            w = want(kind, body)
            assert attach(cro, kind, body) == w, (it, kind, body, w)
            tally["mismatch" if "Original error" in w else "ok"] += 1
    assert tally["ok"] > 400 and tally["mismatch"] > 400, tally
