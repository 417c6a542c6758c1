"""The host-side parsers (exec output, fabric-manager HTTP bodies, Go-JSON codec) under
AddressSanitizer + UndefinedBehaviorSanitizer with a deterministic mutation fuzzer
(tests/host_fuzz.cpp): no out-of-bounds read, overflow, leak or abort on arbitrary bytes."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "composable-resource-operator_b200", "csrc")
HOST_SOURCES = ["gojson.cpp", "identity.cpp", "reconcile.cpp", "detach.cpp", "fabric.cpp", "provider.cpp", "nodes.cpp", "gpus.cpp", "gpus_local.cpp", "gotypes.cpp", "nvml_ops.cpp"]


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ unavailable")
def test_parsers_survive_mutation_fuzz_under_asan_ubsan(tmp_path):
    exe = str(tmp_path / "host_fuzz")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
           os.path.join(ROOT, "tests", "host_fuzz.cpp")] + [os.path.join(CSRC, s) for s in HOST_SOURCES] + ["-o", exe, "-ldl"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "asan" in build.stderr.lower():
        pytest.skip("libasan not installed: " + build.stderr[-200:])
    assert build.returncode == 0, build.stderr[-2000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    run = subprocess.run([exe, "20000"], capture_output=True, text=True, env=env, timeout=300)
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "host fuzz ok: 20000 inputs" in run.stdout
    assert "runtime error" not in run.stderr and "AddressSanitizer" not in run.stderr


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ unavailable")
def test_concurrent_callers_under_tsan(tmp_path):
    """Eight threads drive the same host functions at once under ThreadSanitizer: the parsers, clients and node-side flows
    keep no shared mutable state (the Go host calls them from whatever OS thread a goroutine happens to be on)."""
    exe = str(tmp_path / "host_fuzz_tsan")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
           os.path.join(ROOT, "tests", "host_fuzz.cpp")] + [os.path.join(CSRC, s) for s in HOST_SOURCES] + ["-o", exe, "-ldl", "-lpthread"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and "tsan" in build.stderr.lower():
        pytest.skip("libtsan not installed: " + build.stderr[-200:])
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe, "4000", "8"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0:report_signal_unsafe=0"))
    assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
    assert "host fuzz ok: 4000 inputs" in run.stdout
    assert "ThreadSanitizer" not in run.stderr, run.stderr[-3000:]
