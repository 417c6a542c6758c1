"""Parity tests proper: the CUDA path (through the C ABI) against the oracle.

Bit-exact bar: the kernels compute 64-bit integer checksums; identity strings
are compared byte for byte with what `nvidia-smi` prints on the same box."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

MASK = (1 << 64) - 1
VARIANTS = [1, 2, 3]          # READ_LDG, READ_TMA, READ_LDG256
COPY_VARIANTS = [1, 2]


@pytest.fixture(scope="module")
def ctx_small(cro):
    with cro.ProbeContext(sweep_bytes=64 << 20, devices=[0], flags=cro.F_VERIFY_COPY, read_sweeps=3, copy_sweeps=2) as c:
        yield c


def test_pattern_words_match_oracle(cro, coracle, ctx_small):
    ctx_small.hbm_fill(0)
    seed = ctx_small.seed(0)
    n_words = (64 << 20) // 8
    for first, n in ((0, 4096), (n_words - 1024, 1024), (123457, 999)):
        got = ctx_small.read_words(0, first, n)
        assert got == [coracle.pattern_word(seed, first + i) for i in range(n)]


@pytest.mark.parametrize("variant", VARIANTS)
def test_read_checksum_matches_oracle(cro, coracle, ctx_small, variant):
    s = ctx_small.hbm_read_checksum(0, variant)
    assert (s.checksum_xor, s.checksum_sum) == coracle.checksum(ctx_small.seed(0), 0, (64 << 20) // 8)
    assert s.variant == variant and s.bytes == 64 << 20 and s.ns > 0


def test_expected_kernel_matches_oracle(coracle, ctx_small):
    s = ctx_small.hbm_expected_checksum(0)
    assert (s.checksum_xor, s.checksum_sum) == coracle.checksum(ctx_small.seed(0), 0, (64 << 20) // 8)


@pytest.mark.parametrize("cv", COPY_VARIANTS)
def test_copy_round_trip(cro, coracle, ctx_small, cv):
    want = coracle.checksum(ctx_small.seed(0), 0, (64 << 20) // 8)
    ctx_small.hbm_fill(0)
    c = ctx_small.hbm_copy(0, cv)
    assert c.bytes == 2 * (64 << 20) and c.variant == cv
    for rv in VARIANTS:
        d = ctx_small.hbm_read_checksum(0, rv, dst=True)
        assert (d.checksum_xor, d.checksum_sum) == want, (cv, rv)
    # destination words themselves, not only their checksum
    n_words = (64 << 20) // 8
    assert ctx_small.read_words(0, n_words + 77, 64) == ctx_small.read_words(0, 77, 64)


# ragged and tiny sizes: not a multiple of any tile; the smallest legal sweep is 16 bytes
@pytest.mark.parametrize("nbytes", [16, 32, 4096 + 16, 65536 - 16, 1 << 20, (1 << 20) + 48, 3 * (1 << 20) + 16 * 7,
                                    (32 << 20) + 32784])
def test_ragged_sizes(cro, coracle, nbytes):
    with cro.ProbeContext(sweep_bytes=nbytes, devices=[0], flags=cro.F_VERIFY_COPY, read_sweeps=1, copy_sweeps=1,
                          seed_base=0x1234500000000000) as c:
        want = coracle.checksum(c.seed(0), 0, nbytes // 8)
        for rv in VARIANTS:
            s = c.hbm_read_checksum(0, rv)
            assert (s.checksum_xor, s.checksum_sum) == want, (nbytes, rv)
        for cv in COPY_VARIANTS:
            c.hbm_copy(0, cv)
            d = c.hbm_read_checksum(0, 1, dst=True)
            assert (d.checksum_xor, d.checksum_sum) == want, (nbytes, cv)
        e = c.hbm_expected_checksum(0)
        assert (e.checksum_xor, e.checksum_sum) == want


def test_fault_is_detected_and_located(cro, coracle, ctx_small):
    """A single flipped bit anywhere must change the checksum by exactly that bit; the probe reports it."""
    ctx_small.hbm_fill(0)
    seed, n_words = ctx_small.seed(0), (64 << 20) // 8
    clean = coracle.checksum(seed, 0, n_words)
    for word, bit in ((0, 0), (n_words - 1, 63), (n_words // 3, 17)):
        ctx_small.inject_fault(0, word, 1 << bit)
        for rv in VARIANTS:
            s = ctx_small.hbm_read_checksum(0, rv)
            assert s.checksum_xor == clean[0] ^ (1 << bit), (word, bit, rv)
            assert s.checksum_sum != clean[1]
        ctx_small.inject_fault(0, word, 1 << bit)          # undo
    s = ctx_small.hbm_read_checksum(0, 1)
    assert (s.checksum_xor, s.checksum_sum) == clean
    # the full probe refills, so it passes; corrupt after fill is caught by hbm_read sweeps
    r = ctx_small.probe_device(0)
    assert r.status == 0 and (r.checksum_xor, r.checksum_sum) == clean == (r.expect_xor, r.expect_sum)
    assert (r.copy_checksum_xor, r.copy_checksum_sum) == clean


def test_probe_result_fields(cro, coracle, ctx_small):
    r = ctx_small.probe_device(0)
    d = ctx_small.enumerate()[0]
    assert r.abi_version == 1 and r.status == 0 and r.world == 1 and r.rank == 0
    assert r.gpu_uuid == d.gpu_uuid and r.pci_bus_id == d.pci_bus_id and r.device_minor == d.device_minor
    assert r.sweep_bytes == 64 << 20 and r.read_sweeps == 3 and r.copy_sweeps == 2
    assert r.seed == (0x00C0FFEE00000000 | max(d.device_minor, 0)) or d.device_minor < 0
    assert 0 < r.read_best_ns <= r.read_median_ns and 0 < r.copy_best_ns <= r.copy_median_ns and r.fill_ns > 0
    assert r.sm_count == 148


def test_full_size_probe_matches_oracle(cro, coracle):
    """BASELINE config 2: S = 4 GiB.  The C oracle recomputes the closed form with all host threads."""
    S = 4 << 30
    with cro.ProbeContext(sweep_bytes=S, devices=[0], flags=cro.F_VERIFY_COPY) as c:
        r = c.probe_device(0)
        want = coracle.checksum(r.seed, 0, S // 8, threads=os.cpu_count() or 1)
        assert (r.checksum_xor, r.checksum_sum) == want
        assert (r.copy_checksum_xor, r.copy_checksum_sum) == want
        assert (r.expect_xor, r.expect_sum) == want
        # size-independent property: checksum of the whole == combination of the halves' closed forms
        a = coracle.checksum(r.seed, 0, S // 16, threads=os.cpu_count() or 1)
        x2, s2 = want[0] ^ a[0], (want[1] - a[1]) & MASK
        assert (x2, s2) == coracle.checksum(r.seed, S // 16, S // 16)
        for rv in VARIANTS:
            s = c.hbm_read_checksum(0, rv)
            assert (s.checksum_xor, s.checksum_sum) == want


def test_identity_strings_match_nvidia_smi(cro):
    """cro_emit_csv must print what the reference's exec of nvidia-smi prints (gpus.go:886)."""
    smi = shutil.which("nvidia-smi")
    if not smi:
        pytest.skip("nvidia-smi not on this box")
    with cro.ProbeContext(sweep_bytes=1 << 20, flags=cro.F_LAZY_ALLOC) as c:
        devs = c.enumerate()
        def smi_csv(q):
            return subprocess.run([smi, "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True)
        for q in ("gpu_uuid", "gpu_uuid,pci.bus_id", "index,gpu_uuid,pci.bus_id,name"):
            want = smi_csv(q)
            assert want.returncode == 0, want.stdout + want.stderr
            assert cro.emit_csv(devs, q) == want.stdout, q
        # The reference's 3-field query (gpus.go:216-218).  Some nvidia-smi builds (driver 580 here) reject
        # `device_minor` ("not a valid field to query") — then the minor is pinned through NVML's
        # minor_number spelling if the build has it, and through /proc below.
        want = smi_csv("device_minor,gpu_uuid,pci.bus_id")
        if want.returncode == 0:
            assert cro.emit_csv(devs, "device_minor,gpu_uuid,pci.bus_id") == want.stdout
        else:
            assert "not a valid field" in want.stdout + want.stderr
            alt = smi_csv("minor_number,gpu_uuid,pci.bus_id")
            if alt.returncode == 0:
                assert cro.emit_csv(devs, "minor_number,gpu_uuid,pci.bus_id") == alt.stdout
        # /proc flavour (gpus.go:1017-1037), when the driver exposes it in this container
        base = "/proc/driver/nvidia/gpus"
        if os.path.isdir(base):
            lines = ""
            for name in sorted(os.listdir(base)):
                p = os.path.join(base, name, "information")
                if os.path.isfile(p):
                    lines += cro.proc_information_to_line(open(p).read())
            rc, js = cro.getGPUInfoFromProcOutput(lines, "", None, "device_minor,gpu_uuid,pci.bus_id")
            assert rc == 0
            import json
            by_uuid = {m["gpu_uuid"]: m for m in json.loads(js)}
            for d in devs:
                m = by_uuid[d.gpu_uuid.decode()]
                assert m["device_minor"] == str(d.device_minor)
                assert cro.normalize(0, m["pci.bus_id"]).endswith(cro.normalize(2, d.pci_bus_id.decode()))


def test_reconcile_attach_live(cro, oracle):
    import __graft_entry__ as g
    with cro.ProbeContext(sweep_bytes=32 << 20, devices=[0], read_sweeps=1, copy_sweeps=1) as c:
        uuid = c.enumerate()[0].gpu_uuid.decode()
        base = {"name": "cr-0", "spec": {"type": "gpu", "model": "NVIDIA-B200", "target_node": "worker-0"},
                "status": {"state": "Attaching"}, "device_resource_type": "DEVICE_PLUGIN", "probe": True}
        out = cro.reconcile_attach(c, dict(base, provider={"device_id": uuid, "cdi_device_id": "res-0-0"}))
        assert g.json_status(out) == oracle.emit_status("Online", "", uuid, "res-0-0")
        assert out["probe"]["cohdi.io/probe-status"] == "ok" and out["probe"]["cohdi.io/probe-device-id"] == uuid
        # a device the fabric promised but the node does not have: stays Attaching, 30 s requeue
        out = cro.reconcile_attach(c, dict(base, provider={"device_id": "GPU-00000000-dead-beef-0000-000000000000", "cdi_device_id": "r"}))
        assert g.json_status(out) == oracle.emit_status("Attaching", "", "GPU-00000000-dead-beef-0000-000000000000", "r")
        assert out["requeue_after_s"] == 30 and "probe" not in out


def test_launch_count_is_kernels(cro):
    with cro.ProbeContext(sweep_bytes=16 << 20, devices=[0], read_sweeps=4, copy_sweeps=3) as c:
        c.probe_device(0)
        first = c.launch_count()
        assert first == 1 + 1 + 4 + 3          # expected-checksum + fill + reads + copies
        c.probe_device(0)
        assert c.launch_count() - first == 1 + 4 + 3   # the closed form is cached per device


def test_probe_all_on_a_single_device(cro, coracle):
    """A one-GPU node: no NVLink rounds, no NCCL (nothing to gather from), same result as the per-device probe."""
    S = 64 << 20
    with cro.ProbeContext(sweep_bytes=S, devices=[0], read_sweeps=2, copy_sweeps=1) as c:
        res = c.probe_all()
        assert len(res) == 1
        r = res[0]
        assert r.status == 0 and r.rank == 0 and r.world == 1
        assert (r.checksum_xor, r.checksum_sum) == coracle.checksum(r.seed, 0, S // 8)
        assert all(x == 0 for x in r.p2p_read_ns) and all(x == 0 for x in r.p2p_write_ns)
        one = c.probe_device(0)
        assert (one.checksum_xor, one.checksum_sum, one.gpu_uuid) == (r.checksum_xor, r.checksum_sum, r.gpu_uuid)


def test_multi_device_probe_all(cro, coracle):
    import ctypes
    with cro.ProbeContext(sweep_bytes=256 << 20, p2p_bytes=64 << 20, read_sweeps=2, copy_sweeps=1, latency_hops=2048) as c:
        n = c.device_count()
        if n < 2:
            pytest.skip("single-GPU box")
        res = c.probe_all()
        assert len(res) == n
        for i, r in enumerate(res):
            assert r.status == 0 and r.rank == i and r.world == n
            assert (r.checksum_xor, r.checksum_sum) == coracle.checksum(r.seed, 0, (256 << 20) // 8)
            for j in range(min(n, 8)):
                if j == i or not r.p2p_access[j]:
                    continue
                assert r.p2p_read_ns[j] > 0 and r.p2p_latency_ns_x16[j] > 0 and r.p2p_write_ns[j] > 0
                assert r.p2p_checksum_xor[j] == coracle.checksum(res[j].seed, 0, (64 << 20) // 8)[0]


def test_peer_push_lands_the_pushers_pattern(cro, coracle):
    """The push leg writes a's pattern prefix into the scratch half of b over NVLink; afterwards that half
    must hold exactly a's words (read back through the C ABI and compared with the oracle's generator)."""
    S, P = 64 << 20, 16 << 20
    with cro.ProbeContext(sweep_bytes=S, p2p_bytes=P, read_sweeps=1, copy_sweeps=1, latency_hops=256, flags=cro.F_SKIP_COPY) as c:
        n = c.device_count()
        if n < 2:
            pytest.skip("single-GPU box")
        res = c.probe_all()
        assert all(r.status == 0 for r in res)
        # the LAST round of the 1-factorisation pairs each device with a known partner: find it by content
        seeds = [r.seed for r in res]
        for b in range(n):
            words = c.read_words(b, S // 8, 4)                  # first words of b's scratch half
            owners = [a for a in range(n) if a != b and list(words) == [coracle.pattern_word(seeds[a], i) for i in range(4)]]
            assert len(owners) == 1, (b, words)
            tail = c.read_words(b, S // 8 + P // 8 - 4, 4)       # ...and the last words of the pushed prefix
            assert list(tail) == [coracle.pattern_word(seeds[owners[0]], P // 8 - 4 + i) for i in range(4)]


def test_oom_fails_loudly_or_degrades(cro, coracle):
    """A sweep region that does not fit (2*S = 192 GiB > 180 GB): CRO_ERR_OOM by default; with
    CRO_F_DEGRADE_ON_OOM the probe halves S until it fits and says so in the result."""
    S = 96 << 30
    import torch
    torch.cuda.init()
    free_before = torch.cuda.mem_get_info(0)[0]
    for _ in range(3):                                   # a failed init must release what it had already built
        with pytest.raises(cro.ProbeError) as e:
            cro.ProbeContext(sweep_bytes=S, devices=[0])
        assert e.value.code == cro.ERR_OOM
        assert "cudaMalloc" in str(e.value) and "asked for" in str(e.value)      # cro_last_error(NULL) carries the reason
    assert free_before - torch.cuda.mem_get_info(0)[0] < (64 << 20)
    with cro.ProbeContext(sweep_bytes=S, devices=[0], flags=cro.F_DEGRADE_ON_OOM, read_sweeps=1, copy_sweeps=1) as c:
        r = c.probe_device(0)
        assert r.status == 0 and r.sweep_bytes == 48 << 30
        assert (r.checksum_xor, r.checksum_sum) == coracle.checksum(r.seed, 0, r.sweep_bytes // 8, threads=os.cpu_count() or 1)


def test_concurrent_callers_are_serialised_per_device(cro, coracle):
    """Reconciles for different CRs may probe the same GPU from different OS threads (cgo migrates
    goroutines): every entry point takes the device mutex and calls cudaSetDevice itself."""
    import threading
    with cro.ProbeContext(sweep_bytes=32 << 20, devices=[0], read_sweeps=2, copy_sweeps=1) as c:
        want = coracle.checksum(c.seed(0), 0, (32 << 20) // 8)
        errors, results = [], []

        def worker(k):
            try:
                for i in range(4):
                    if (k + i) % 3 == 0:
                        s = c.hbm_read_checksum(0, 1 + (k + i) % 3)
                        results.append((s.checksum_xor, s.checksum_sum))
                    elif (k + i) % 3 == 1:
                        r = c.probe_device(0)
                        results.append((r.checksum_xor, r.checksum_sum))
                    else:
                        out = cro.reconcile_attach(c, {"status": {"state": "Attaching"}, "probe": True, "spec": {"type": "gpu", "model": "m", "target_node": "n"},
                                                       "provider": {"device_id": c.enumerate()[0].gpu_uuid.decode(), "cdi_device_id": "r"}})
                        assert out["status"]["state"] == "Online"
            except Exception as e:   # noqa: BLE001
                errors.append(repr(e))
        ts = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert errors == [] and results and all(r == want for r in results)


def test_cli_helper_process(cro):
    """croprobe-cli: the fresh-process form (a hot-plugged GPU is invisible to an already initialised CUDA process)."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cli = os.path.join(root, "composable-resource-operator_b200", "croprobe-cli")
    smi = shutil.which("nvidia-smi")
    if smi:
        want = subprocess.run([smi, "--query-gpu=gpu_uuid,pci.bus_id", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
        got = subprocess.run([cli, "csv", "gpu_uuid,pci.bus_id"], capture_output=True, text=True)
        assert got.returncode == 0 and got.stdout == want
    devs = json.loads(subprocess.run([cli, "enumerate"], capture_output=True, text=True).stdout)
    uuid = devs[0]["gpu_uuid"]
    out = subprocess.run([cli, "probe", uuid, "256"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    ann = json.loads(out.stdout)
    assert ann["cohdi.io/probe-status"] == "ok" and ann["cohdi.io/probe-device-id"] == uuid
    assert subprocess.run([cli, "probe", "GPU-00000000-dead-beef-0000-000000000000"], capture_output=True).returncode == 3
    cold = json.loads(subprocess.run([cli, "cold", "0", "4096"], capture_output=True, text=True).stdout)
    assert cold["status"] == 0 and cold["cold_total_s"] > cold["warm_probe_s"] > 0
    print("cold vs warm:", cold)


def test_c_harness_on_gpu(cro):
    """The plain-C caller (what cgo compiles to) runs a probe + emit through the same ABI."""
    from test_abi import build_c_harness
    out = subprocess.run([build_c_harness(), "gpu"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "gpu ok: GPU-" in out.stdout and "cohdi.io/probe-status" in out.stdout


def test_async_probe_begin_end(cro, coracle):
    """cro_probe_begin / cro_probe_end: same result as the synchronous probe; a sweep in between drains it."""
    with cro.ProbeContext(sweep_bytes=32 << 20, devices=[0], read_sweeps=2, copy_sweeps=1) as c:
        want = coracle.checksum(c.seed(0), 0, (32 << 20) // 8)
        c.probe_begin(0)
        c.probe_begin(0)                       # second begin is a no-op
        r = c.probe_end(0)
        assert r.status == 0 and (r.checksum_xor, r.checksum_sum) == want
        r2 = c.probe_end(0)                    # end without begin probes synchronously
        assert r2.status == 0 and (r2.checksum_xor, r2.checksum_sum) == want
        c.probe_begin(0)
        s = c.hbm_read_checksum(0, 1)          # another op first drains the in-flight probe
        assert (s.checksum_xor, s.checksum_sum) == want
        r3 = c.probe_end(0)
        assert r3.status == 0 and r3.read_best_ns > 0


def test_storm_and_churn_with_live_probe(cro):
    """BASELINE configs 4 / 5 in miniature with the CUDA probe in the attach slot (all GPUs of the box)."""
    import random
    with cro.ProbeContext(sweep_bytes=64 << 20, read_sweeps=2, copy_sweeps=1) as ctx:
        n = ctx.device_count()
        uuids = [d.gpu_uuid.decode() for d in ctx.enumerate()]
        with cro.Cluster({"nodes": ["worker-%d" % i for i in range(n)], "probe": True}, ctx) as c:
            rng = random.Random(1)
            sizes = {}
            for i in range(24):
                sizes["req-%02d" % i] = rng.randint(1, 3)
                assert c.apply("req-%02d" % i, {"type": "gpu", "model": "NVIDIA-B200-%d" % (i // n), "size": sizes["req-%02d" % i],
                                                "target_node": "worker-%d" % (i % n)}) == ""
            st = c.run()
            assert st["requests_running"] == 24 and st["reconcile_errors"] == 0 and st["probe_failures"] == 0
            assert st["probes"] == sum(sizes.values())          # every attach was probed exactly once
            d = c.dump()
            for name, req in d["requests"].items():
                node = int(req["spec"]["target_node"].split("-")[1])
                assert all(cs["state"] == "Online" and cs["device_id"] == uuids[node] for cs in req["status"]["resources"].values())
            for name in sizes:
                c.delete(name)
            c.run()
            d = c.dump()
            assert d["requests"] == {} and d["resources"] == {}
