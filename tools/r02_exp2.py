"""Where does a reconcile's wall time go on a multi-GPU node?  One process, context over device 0 only."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cro = importlib.import_module("composable-resource-operator_b200")
S = 4 << 30


def t(fn, n):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return round((time.perf_counter() - t0) / n * 1e3, 4)


with cro.ProbeContext(sweep_bytes=S, devices=[int(sys.argv[1]) if len(sys.argv) > 1 else 0]) as c:
    uuid = c.own_devices()[0].gpu_uuid.decode()
    req = {"name": "cr", "spec": {"type": "gpu", "model": "m", "target_node": "n"}, "status": {"state": "Attaching"}, "probe": True,
           "device_resource_type": "DEVICE_PLUGIN", "provider": {"device_id": uuid, "cdi_device_id": "r"}}
    out = {"enumerate_ms": t(lambda: c.enumerate(), 50), "probe_device_ms": t(lambda: c.probe_device(0), 10),
           "probe_uuid_ms": t(lambda: cro.probe_uuid(c, uuid), 10), "reconcile_ms": t(lambda: cro.reconcile_attach(c, req), 10),
           "node": [(d.gpu_uuid.decode(), d.flags, d.dev_index) for d in c.enumerate()]}
    ts = c.sweep_times(0)
    out["sweeps_event_gbs"] = [[x.kind, x.index, round(x.bytes / x.event_ns, 1), round(x.bytes / x.timer_ns, 1)] for x in ts]
    print(json.dumps(out))
