"""JSON nested deeper than the tree builder recurses (512) but within encoding/json's own limit (10000): Go decodes such a
reply — the deep part can only sit in an unknown field, a json.RawMessage or a map[string]any — so it must decode here too;
beyond 10000 levels the error is Go's scanner's.  (Expected strings written out: the Python oracle's json.loads gives up at
its own recursion limit long before.)"""
import importlib
import json

import pytest

from test_error_body_decode import attach


@pytest.fixture(scope="module")
def cro():
    return importlib.import_module("composable-resource-operator_b200")


def deep(n, core='{"k":"]}\\\\\\"["}'):
    return "[" * n + core + "]" * n


def test_deep_value_in_an_unknown_field_is_skipped(cro):
    body = '{"junk":%s,"status":404,"detail":{"code":"E02","message":"machine not found","more":%s}}' % (deep(3000), deep(700))
    assert attach(cro, "CM", body) == \
        "failed to process CM get request. http returned status: '404', cm return code: 'E02', error message: 'machine not found'"


def test_deep_raw_message_is_reported_as_its_own_text(cro):
    msg = deep(800)
    got = attach(cro, "FM", '{"detail":{"code":"E9","message": %s }}' % msg)
    assert got == "failed to process FM scaleup request. FM returned code: 'E9', error message: '%s'" % msg.replace('\\\\\\"', '\\\\\\"')
    assert len(got) > 1600


def test_deep_map_content_is_accepted(cro):
    assert attach(cro, "FM", '{"detail":{"code":"E1","message":"m","data":{"x":%s}}}' % deep(5000)) == \
        "failed to process FM scaleup request. FM returned code: 'E1', error message: 'm'"


def test_deep_value_where_a_typed_field_is_expected_still_mismatches(cro):
    got = attach(cro, "CM", '{"status":%s}' % deep(600))
    assert got == ("failed to unmarshal CM get error response body into errBody. Original error: "
                   "json: cannot unmarshal array into Go struct field ErrorBody.status of type int")


def test_beyond_go_s_limit_is_go_s_error(cro):
    got = attach(cro, "CM", "[" * 10001 + "]" * 10001)
    assert got == ("failed to unmarshal CM get error response body into errBody. Original error: "
                   "invalid character '[' exceeded max depth")
    ok = attach(cro, "CM", "[" * 10000 + "]" * 10000)
    assert ok == ("failed to unmarshal CM get error response body into errBody. Original error: "
                  "json: cannot unmarshal array into Go value of type api.ErrorBody")


def test_harness_request_itself_may_be_deep(cro):
    out = cro.token_from_reply({"status": 200, "body": json.dumps({"access_token": "a.b"}), "extra": json.loads("[" * 400 + "]" * 400)})
    assert out["error"] == "invalid access token: a.b"
