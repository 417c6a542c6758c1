"""csrc/gpus.cpp vs oracle/node_ops.py on random scripted clusters: every flavour (RKE2 / OCP x DRA / DEVICE_PLUGIN),
missing pods, unset ClusterPolicy, failing and garbage-printing commands.  Compared: the reconcile error, and the
exact sequence of pod-execs (pod, container, URL query)."""
import random

import node_ops as no

DEV = "GPU-7cc45b7b-2a6d-f0ac-1b02-6f8de09e1a6c"
OUTS = ["", "\n", "No devices were found\n", "0, %s, 00000000:1F:00.0\n" % DEV, "0, %s, 00000000:1F:00.0\n1, GPU-x, 00000000:2F:00.0\n" % DEV,
        "0,%s,0000:1f:00.0\n" % DEV, "0,%s,0000:1f:00.0\n1,GPU-x,0000:2f:00.0\n" % DEV, DEV, "GPU-x\n" + DEV + "\n", DEV + ", python3\n",
        "GPU-x, trainer\n", "currently: draining\n", "GPU 0000:1F:00.0 is currently: not draining.\n", "garbage", "a,b", "true\n",
        "Module Size\nnvidia_drm 1 0\nnvidia_uvm 2 0\n", "nvidia_uvm 2 0\n", "nvidia 3 0\n", "nvidia-persist\n", "12 python3, 14 x"]
NEEDLES = [{"escape": "--query-gpu=gpu_uuid"}, {"escape": "--query-gpu=device_minor,gpu_uuid,pci.bus_id"},
           {"escape": "--query-compute-apps=gpu_uuid,process_name"}, {"escape": "/proc/driver/nvidia/gpus"}, {"escape": "TARGET_FILE"},
           {"escape": "TARGET="}, {"literal": "command=-pm&command=0"}, {"literal": "command=-m&command=1"}, {"literal": "command=-r"},
           {"literal": "command=-q"}, {"escape": "/usr/sbin/lsmod"}, {"escape": "/usr/sbin/modprobe"}, {"escape": "/usr/bin/tee"},
           {"escape": "/run/nvidia/driver/dev/nvidia"}, {"escape": "/dev/nvidia"}, None]
PODS = [{"namespace": "nvidia-gpu-operator", "name": "nvidia-driver-daemonset-a", "node": "worker-0",
         "labels": {"app.kubernetes.io/component": "nvidia-driver"}, "containers": ["nvidia-driver-ctr"]},
        {"namespace": "nvidia-gpu-operator", "name": "nvidia-driver-daemonset-b", "node": "worker-1",
         "labels": {"app.kubernetes.io/component": "nvidia-driver"}, "containers": ["other ctr"]},
        {"namespace": "nvidia-dra-driver-gpu", "name": "nvidia-dra-driver-gpu-kubelet-plugin-x", "node": "worker-0",
         "labels": {"app.kubernetes.io/name": "nvidia-dra-driver-gpu"}, "containers": ["compute-domains"]},
        {"namespace": "cro", "name": "cro-node-agent-q", "node": "worker-0", "labels": {"app": "cro-node-agent"}, "containers": ["agent"]},
        {"namespace": "cro", "name": "unrelated", "node": "worker-0", "labels": {"app": "cro-node-agent"}, "containers": []}]


def test_gpus_cpp_vs_python_restatement(cro):
    rng = random.Random(424242)
    seen = set()
    for it in range(1500):
        pods = [p for p in PODS if rng.random() < 0.8]
        rng.shuffle(pods)
        rules = []
        for n in rng.sample(NEEDLES[:-1], rng.randrange(4, len(NEEDLES))):
            rule = {"needle": n, "stdout": rng.choice(OUTS) if rng.random() < 0.8 else "", "stderr": "" if rng.random() < 0.9 else "boom"}
            if rng.random() < 0.04:
                rule["exec_err"] = "command terminated with exit code 1"
            rules.append(rule)
        rules.append({"needle": None, "stdout": "", "stderr": "" if rng.random() < 0.7 else "this error should be reported"})
        policy = rng.choice([None, {}, {"driver_enabled": True}, {"driver_enabled": True}, {"driver_enabled": False}])
        cluster = {"cluster_policy": policy, "pods": pods, "exec": rules}
        dtype = rng.choice(["DRA", "DEVICE_PLUGIN"])
        state = rng.choice(["Attaching", "Detaching", "Detaching"])
        slices = rng.choice([[], [{"devices": [{"attributes": {"uuid": DEV}}]}]])
        req = {"name": "cr", "spec": {"type": "gpu", "model": "m", "target_node": "worker-0", "force_detach": rng.random() < 0.2},
               "status": {"state": state, "device_id": DEV, "cdi_device_id": "res"}, "deleting": state == "Detaching",
               "device_resource_type": dtype, "probe": False, "provider": {}, "resource_slices": slices, "cluster": cluster}
        out = cro.reconcile_attach(None, req)

        c = no.Cluster(cluster, slices)
        status_error = ""

        def is_panic(e):                 # a Go run-time panic: unwinds at once, no status write, "panic: ... [recovered]"
            return e.startswith("runtime error: ")

        def rec(e):
            return "panic: %s [recovered]" % e if is_panic(e) else e
        want_err = ""
        if state == "Attaching":         # composableresource_controller.go:239-286 with the ids already present
            stop = False
            if dtype == "DEVICE_PLUGIN":
                e = no.check_no_gpu_loads(c, "worker-0", None)                      # an ERROR is only logged ...
                if is_panic(e):                                                     # ... a panic is not an error
                    want_err, stop = rec(e), True
            else:
                err = no.run_nvidia_smi(c, "worker-0")
                if is_panic(err):
                    want_err, stop = rec(err), True
                elif err:
                    status_error = err                                              # recorded, flow continues
            if not stop:
                vis, err = no.check_gpu_visible(c, dtype, "worker-0", DEV)
                want_err = rec(err)
                if not err and vis:
                    status_error = ""
        else:                            # :320-407 up to the point where the canned provider says "removed"
            if not req["spec"]["force_detach"]:
                want_err = rec(no.check_no_gpu_loads(c, "worker-0", None if dtype == "DEVICE_PLUGIN" else DEV))
            if not want_err:
                want_err = rec(no.drain_gpu(c, "worker-0", DEV, dtype))
            if not want_err:
                vis, e = no.check_gpu_visible(c, dtype, "worker-0", DEV)
                want_err = rec(e)
        assert out["error"] == want_err, (it, req, out["error"], want_err)
        got = [(x["pod"], x["container"], x["query"], x["kind"], x["detached"]) for x in out["exec_log"]]
        want = [(x["pod"], x["container"], x["query"], x["kind"], x["detached"]) for x in c.log]
        assert got == want, (it, req, got, want)
        assert out["slept_s"] == c.slept
        if state == "Attaching" and not want_err:
            assert out["status"].get("error", "") == status_error, (it, out["status"], status_error)
        seen.add((state, dtype, None if policy is None else policy.get("driver_enabled", "unset"), bool(want_err), len(want) > 6))
    assert len(seen) >= 24, sorted(map(str, seen))       # the generator reaches every flavour, failing and succeeding
