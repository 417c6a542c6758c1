"""The id_manager's answer -> CachedToken.GetToken result (internal/cdi/fti/token.go:72-175): C++ (fabric::TokenFromReply behind
the harness key fabric.token) against the Python restatement, on the reference's six scenarios, hand cases and a mutation fuzz."""
import base64
import importlib
import json
import os
import random
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
import fabric_clients as fc  # noqa: E402

NOW = "2025-06-01T12:00:00Z"
NOW_UNIX = 1748779200


@pytest.fixture(scope="module")
def cro():
    return importlib.import_module("composable-resource-operator_b200")


def b64(raw):
    return base64.urlsafe_b64encode(raw).rstrip(b"=").decode()


def jwt(claims):
    return "h." + b64(json.dumps(claims).encode()) + ".s"


def got(cro, token):
    out = cro.token_from_reply(token)
    return ("unable to rotate token: " + out["error"]) if out["error"] else "", out["expiry"]


def want(token):
    f = fc.Fabric({"token": token})
    f.now = NOW_UNIX
    return f.token()


CASES = [
    ({"secret_error": 'secrets "credentials" not found'}, 'unable to rotate token: secrets "credentials" not found'),
    ({"transport_error": 'Post "https://x/token": dial tcp: connection refused'},
     'unable to rotate token: Post "https://x/token": dial tcp: connection refused'),
    ({"status": 401, "body": '{"error":"invalid_grant"}'}, 'unable to rotate token: http returned code: 401, response body: {"error":"invalid_grant"}'),
    ({"status": 200, "body": "<html>"}, "unable to rotate token: failed to read id_manager response body into Token: invalid character '<' looking for beginning of value"),
    ({"status": 200, "body": ""}, "unable to rotate token: failed to read id_manager response body into Token: unexpected end of JSON input"),
    ({"status": 200, "body": "null"}, "unable to rotate token: invalid access token: "),
    ({"status": 200, "body": "{}"}, "unable to rotate token: invalid access token: "),
    ({"status": 200, "body": '{"access_token":"a.b"}'}, "unable to rotate token: invalid access token: a.b"),
    ({"status": 200, "body": '{"access_token":"a.b.c.d"}'}, "unable to rotate token: invalid access token: a.b.c.d"),
    ({"status": 200, "body": '{"access_token":"h.%%%%.s"}'}, "unable to rotate token: failed to decode id_manager payload: illegal base64 data at input byte 0"),
    ({"status": 200, "body": '{"access_token":"h.e30=.s"}'}, "unable to rotate token: failed to decode id_manager payload: illegal base64 data at input byte 3"),
    ({"status": 200, "body": '{"access_token":"h.e30+.s"}'}, "unable to rotate token: failed to decode id_manager payload: illegal base64 data at input byte 3"),
    ({"status": 200, "body": '{"access_token":"h.e30ab.s"}'}, "unable to rotate token: failed to decode id_manager payload: illegal base64 data at input byte 4"),
    ({"status": 200, "body": '{"access_token":"h..s"}'}, "unable to rotate token: failed to unmarshal id_manager json: unexpected end of JSON input"),
    ({"status": 200, "body": '{"access_token":"h.%s.s"}' % b64(b"this is not json")},
     "unable to rotate token: failed to unmarshal id_manager json: invalid character 'h' in literal true (expecting 'r')"),
    ({"status": 200, "body": '{"access_token":"h.%s.s"}' % b64(b"[1]")},
     "unable to rotate token: failed to unmarshal id_manager json: json: cannot unmarshal array into Go value of type fti.accessToken"),
    # field-level mismatches (go1.24 UnmarshalTypeError wording; unpinned by the reference)
    ({"status": 200, "body": '{"access_token":5}'},
     "unable to rotate token: failed to read id_manager response body into Token: json: cannot unmarshal number into Go struct field token.access_token of type string"),
    ({"status": 200, "body": '{"expires_in":1.5,"access_token":[]}'},
     "unable to rotate token: failed to read id_manager response body into Token: json: cannot unmarshal number 1.5 into Go struct field token.expires_in of type int64"),
    ({"status": 200, "body": '{"Not-Before-Policy":"x"}'},
     "unable to rotate token: failed to read id_manager response body into Token: json: cannot unmarshal string into Go struct field token.not-before-policy of type int64"),
    ({"status": 200, "body": '{"access_token":"h.%s.s"}' % b64(b'{"exp":99999999999999999999}')},
     "unable to rotate token: failed to unmarshal id_manager json: json: cannot unmarshal number 99999999999999999999 into Go struct field accessToken.exp of type int64"),
    ({"status": 200, "body": '{"access_token":"%s","access_token":null}' % jwt({"exp": NOW_UNIX + 3600})}, ""),   # null leaves the field
    ({"status": 200, "body": '{"ACCESS_TOKEN":"%s"}' % jwt({"exp": NOW_UNIX + 3600})}, ""),      # keys fold
    ({"status": 200, "body": '{"access_token":"%s"}' % jwt({"exp": NOW_UNIX + 3600})}, ""),
    ({"status": 200, "body": '{"access_token":"%s"}' % jwt({})}, ""),                              # exp 0: accepted, never cached
    ({"status": 200, "body": '{"access_token":"h.e3\\n0.s"}'}, ""),                                  # CR / LF inside base64 are skipped
]


@pytest.mark.parametrize("i", range(len(CASES)))
def test_hand_cases(cro, i):
    token, err = CASES[i]
    assert want(token) == err
    assert got(cro, token)[0] == err
    if not err:
        assert got(cro, token)[1] == fc.token_from_reply(token)[0]


def _attach(cro, token, kind):
    return cro.reconcile_attach(None, {
        "name": "cr", "spec": {"type": "gpu", "model": "m", "target_node": "worker-0"}, "status": {"state": "Attaching"},
        "deleting": False, "probe": False, "now": NOW,
        "env": {"DEVICE_RESOURCE_TYPE": "DRA", "CDI_PROVIDER_TYPE": "FTI_CDI", "FTI_CDI_API_TYPE": kind, "FTI_CDI_TENANT_ID": "t", "FTI_CDI_CLUSTER_ID": "c"},
        "fabric": {"token": token, "objects": {
            "nodes": {"worker-0": {"annotations": {"machine.openshift.io/machine": "ns/m0"}}},
            "metal3machines": {"ns/m0": {"annotations": {"metal3.io/BareMetalHost": "ns/b0"}}},
            "baremetalhosts": {"ns/b0": {"annotations": {"cluster-manager.cdi.io/machine": "mach"}}}},
            "http": [{"path_contains": "", "status": 500, "body": "{}"}]},
        "enumeration": {"stdout": "", "stderr": ""}, "resource_slices": []})


@pytest.mark.parametrize("kind", ["CM", "FM"])
def test_cache_window(cro, kind):
    """A token is reused while expiry - 30 s is ahead (token.go:68,78): CM's AddResource asks twice (client.go:113,
    getMachineInfo :395), so a token about to lapse costs a second id_manager round trip, a fresh one does not."""
    fresh = {"status": 200, "body": '{"access_token":"%s"}' % jwt({"exp": NOW_UNIX + 3600})}
    lapsing = {"status": 200, "body": '{"access_token":"%s"}' % jwt({"exp": NOW_UNIX + 30})}
    a, b = _attach(cro, fresh, kind), _attach(cro, lapsing, kind)
    asks = len(a["fabric_requests"])               # every fabric request is preceded by one GetToken
    assert asks >= 1 and a["token_fetches"] == 1
    assert b["token_fetches"] == len(b["fabric_requests"]) == asks
    f = fc.Fabric({"token": lapsing})
    f.now = NOW_UNIX
    for _ in range(asks):
        assert f.token() == ""
    assert f.token_fetches == asks


def test_fuzz_against_restatement(cro):
    rng = random.Random(20250921)
    seeds = [c[0]["body"] for c in CASES if c[0].get("body")]
    payloads = [b'{"exp":1748782800}', b'{"exp":"x"}', b'{"exp":1.0}', b'{"exp":null}', b'{"exp":-5,"exp":true}', b"this is not json", b"[1]", b"null", b'{"EXP":5}', b"", b"\xff\xfe", b'{"exp":1}x']
    tally = {}
    for it in range(1500):
        r = rng.random()
        if r < 0.45:
            mid = b64(rng.choice(payloads))
            if rng.random() < 0.5 and mid:
                k = rng.randrange(len(mid) + 1)
                mid = mid[:k] + rng.choice(["=", "%", "\n", "\r", "+", "/", "A", "_", "-", "é", " "]) + mid[k + (rng.random() < 0.5):]
            elif rng.random() < 0.3:
                mid = mid[:rng.randrange(len(mid) + 1)]
            body = json.dumps({rng.choice(["access_token", "Access_Token", "access_token", "accessToken"]):
                               rng.choice(["h.", "", "h.x."]) + mid + rng.choice([".s", "", ".s.t"])})
        else:
            s = list(rng.choice(seeds))
            for _ in range(rng.randrange(0, 3)):
                k = rng.randrange(len(s) + 1)
                op = rng.random()
                if op < 0.4 and s:
                    del s[min(k, len(s) - 1)]
                elif op < 0.8:
                    s.insert(k, rng.choice('{}[]":,.\\x01 \n<'))
                else:
                    s[k:k] = list(rng.choice(["null", "true", "1e5", '"access_token"', '"expires_in":7,', '"scope":{},']))
            body = "".join(s)
        token = {"status": rng.choice([200] * 12 + [204, 401, 500]), "body": body}
        w = want(token)
        g = got(cro, token)
        assert g[0] == w, (it, token, g, w)
        if not w:
            assert g[1] == fc.token_from_reply(token)[0], (it, token)
        key = w.split(":")[1].strip()[:28] if w else "ok"
        tally[key] = tally.get(key, 0) + 1
    assert len(tally) >= 6, tally
    assert tally.get("ok", 0) >= 20, tally
