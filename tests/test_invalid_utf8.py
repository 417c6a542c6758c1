"""Strings of a fabric reply that are not well-formed UTF-8: encoding/json's unquote() runs utf8.DecodeRune over them, so
every byte that does not start a well-formed sequence becomes one U+FFFD (not one per "maximal subpart", as Python's
errors="replace" does), and a \\uD800 without its partner becomes U+FFFD too.  What json.Marshal then writes is the
replacement character's own three bytes.  Checked through cro_fabric_get_resources with raw bytes against a direct
restatement of utf8.DecodeRune's acceptance table."""
import ctypes
import importlib
import json
import random

import pytest


@pytest.fixture(scope="module")
def cro():
    return importlib.import_module("composable-resource-operator_b200")


def go_decode(raw: bytes) -> str:
    """string(bytes) as Go's unquote sees it."""
    out, i, n = [], 0, len(raw)
    while i < n:
        b0 = raw[i]
        if b0 < 0x80:
            out.append(chr(b0))
            i += 1
            continue
        lo, hi, size = 0x80, 0xBF, 0
        if 0xC2 <= b0 <= 0xDF:
            size = 2
        elif 0xE0 <= b0 <= 0xEF:
            size, lo, hi = 3, (0xA0 if b0 == 0xE0 else 0x80), (0x9F if b0 == 0xED else 0xBF)
        elif 0xF0 <= b0 <= 0xF4:
            size, lo, hi = 4, (0x90 if b0 == 0xF0 else 0x80), (0x8F if b0 == 0xF4 else 0xBF)
        ok = size and i + size <= n and lo <= raw[i + 1] <= hi and all(0x80 <= raw[i + k] <= 0xBF for k in range(2, size))
        if ok:
            out.append(raw[i:i + size].decode("utf-8"))
            i += size
        else:
            out.append("\ufffd")
            i += 1
    return "".join(out)


def device_id_of(cro, payload: bytes) -> str:
    body = b'{"data":{"machines":[{"resources":[{"res_type":"gpu","res_serial_num":"' + payload + b'","res_uuid":"r"}]}]}}'
    buf = ctypes.create_string_buffer(1 << 16)
    n = ctypes.c_size_t(0)
    rc = cro.lib.cro_fabric_get_resources(b"fm", body, b"node", b"mu", buf, len(buf), ctypes.byref(n))
    assert rc == 0, rc
    text = buf.raw[:n.value]
    text.decode("utf-8")                           # what comes out is always well-formed
    assert b"\\ufffd" not in text                  # U+FFFD travels as its own bytes, not as an escape
    return json.loads(text)[0]["device_id"]


CASES = [b"plain", b"\xe2\x82\xac", b"\xff", b"\xe2\x82", b"\xe2\x82D", b"\xed\xa0\x80", b"\xc0\xaf", b"\xc1\xbf", b"\xf4\x90\x80\x80",
         b"\xf0\x9f\x98\x80", b"\xf0\x9f\x98", b"\x80\x80", b"\xe0\x9f\xbf", b"\xe0\xa0\x80", b"\xf5\x80\x80\x80", b"\xef\xbf\xbd"]


@pytest.mark.parametrize("i", range(len(CASES)))
def test_hand_cases(cro, i):
    assert device_id_of(cro, b"A" + CASES[i] + b"Z") == "A" + go_decode(CASES[i]) + "Z"


def test_lone_surrogate_escapes(cro):
    assert device_id_of(cro, b"\\ud800") == "\ufffd"
    assert device_id_of(cro, b"\\ud800\\u0041") == "\ufffdA"
    assert device_id_of(cro, b"\\udc00\\ud83d\\ude00") == "\ufffd\U0001F600"
    assert device_id_of(cro, b"\\ud83d\xff") == "\ufffd\ufffd"


def test_random_bytes(cro):
    rng = random.Random(5)
    alphabet = [0x41, 0x7A, 0x80, 0xBF, 0xC0, 0xC2, 0xDF, 0xE0, 0xE2, 0xED, 0xEF, 0xF0, 0xF4, 0xF5, 0xFF, 0x9F, 0xA0, 0x90, 0x8F, 0x82, 0xAC]
    differs_from_python = 0
    for _ in range(1500):
        raw = bytes(rng.choice(alphabet) for _ in range(rng.randrange(1, 9)))
        want = go_decode(raw)
        assert device_id_of(cro, raw) == want, raw
        differs_from_python += want != raw.decode("utf-8", "replace")
    assert differs_from_python > 50                # the one-per-byte rule is really exercised
