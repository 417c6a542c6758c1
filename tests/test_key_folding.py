"""encoding/json fills a struct field from a key that matches its tag exactly OR case-insensitively
(bytes.EqualFold), keys taken in input order, the last match winning.  A fabric that capitalises a key must
therefore be read the way the Go reference reads it.  C++ reader (csrc/gojson.cpp Value::get) vs the Python
restatement (oracle.GoObj) vs hand-derived expectations."""
import json
import random

DEV, RES, MODEL = "GPU-device00-uuid-temp-0000-000000000000", "GPU-device00-uuid-temp-0000-000000000res", "NVIDIA-A100-PCIE-80GB"


def fm_body(keys):
    res = {keys.get("res_uuid", "res_uuid"): RES, keys.get("res_type", "res_type"): "gpu", keys.get("res_op_status", "res_op_status"): "0",
           keys.get("res_serial_num", "res_serial_num"): DEV,
           keys.get("res_spec", "res_spec"): {"condition": [{"column": "model", "operator": "eq", "value": MODEL}]}}
    return json.dumps({keys.get("data", "data"): {keys.get("machines", "machines"): [{"resources": [res]}]}})


def test_capitalised_and_exotic_keys_still_name_the_field(cro, oracle):
    for keys in ({}, {"res_op_status": "Res_Op_Status"}, {"data": "DATA", "machines": "Machines"}, {"res_serial_num": "RES_SERIAL_NUM"},
                 {"res_spec": "reſ_spec"},            # U+017F LATIN SMALL LETTER LONG S folds to 's'
                 {"res_uuid": "res_uuid"}):
        body = fm_body(keys)
        want = (DEV, RES, "")
        assert oracle.fm_scale_up_response_to_ids(body, "cr", "gpu", MODEL) == want, keys
        assert cro.fm_parse_scale_up_response(body, "cr", "gpu", MODEL) == want, keys
        assert cro.fabric_check_resource("fm", body.replace('"resources"', '"Resources"'), "gpu", MODEL, DEV) == ""


def test_the_last_matching_key_wins_whatever_its_case(cro, oracle):
    base = json.loads(fm_body({}))
    res = base["data"]["machines"][0]["resources"][0]
    inner = json.dumps(res)[1:-1]
    for tail, want_err in ((', "RES_OP_STATUS": "2"', "the FM attached device called by cr is in Critical state in FM"),
                           (', "RES_OP_STATUS": "2", "res_op_status": "0"', ""),
                           (', "res_op_status": "3", "Res_op_status": "1"', "")):
        body = '{"data":{"machines":[{"resources":[{' + inner + tail + '}]}]}}'
        got = cro.fm_parse_scale_up_response(body, "cr", "gpu", MODEL)
        assert got == oracle.fm_scale_up_response_to_ids(body, "cr", "gpu", MODEL)
        assert got[2] == want_err, (tail, got)


def test_random_key_case_fuzz_vs_oracle(cro, oracle):
    rng = random.Random(99)
    from test_cm_provider import cm_machine_data
    cm = cm_machine_data([(DEV, "ADD_COMPLETE", "", RES, "1")])
    fm = fm_body({})
    for _ in range(400):
        kind, body = rng.choice((("cm", cm), ("fm", fm)))
        chars = list(body)
        for _k in range(rng.randrange(1, 6)):              # flip the case of a few letters inside quoted keys and values alike
            i = rng.randrange(len(chars))
            if chars[i].isalpha():
                chars[i] = chars[i].swapcase()
        mutated = "".join(chars)
        try:
            json.loads(mutated)                            # a flipped `null` / `true` is a syntax error: other tests cover those
        except ValueError:
            continue
        assert cro.fabric_check_resource(kind, mutated, "gpu", MODEL, DEV) == oracle.fabric_check_resource(kind, mutated, "gpu", MODEL, DEV), mutated
        assert cro.fabric_get_resources(kind, mutated, "n", "m") == oracle.fabric_get_resources(kind, mutated, "n", "m"), mutated
