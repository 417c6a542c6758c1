"""CPU restatement (TEST INFRASTRUCTURE — only tests/ may import this) of the node-side operations of
internal/utils/gpus.go: pod choice, CheckNoGPULoads (:88-186), DrainGPU (:188-664), RunNvidiaSmi (:666-689),
CheckGPUVisible (:54-86), device taints (:691-766), over the same scripted cluster JSON the C harness takes
("cluster": cluster_policy / pods / exec rules).  Written as one flat interpreter of "steps" rather than the
class hierarchy of csrc/gpus.cpp; tests/test_node_ops_fuzz.py runs both on random clusters.

Pins: the reference's own mock executors (tests/test_node_side_entries.py) pin the OCP flavours; the RKE2
flavours have no reference test — for those this file and gpus.cpp only check each other (parity unpinned)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple
from urllib.parse import quote_plus

import oracle as _o


def _escape(s: str) -> str:          # net/url.QueryEscape == quote_plus with only -_.~ left alone
    return quote_plus(s, safe="-_.~")


def raw_query(argv: List[str], container: str) -> str:
    q = "&".join("command=" + _escape(a) for a in argv)
    return (q + "&" if q else "") + "container=" + _escape(container) + "&stderr=true&stdout=true"


class Cluster:
    def __init__(self, spec: Dict, resource_slices: Optional[List[Dict]] = None):
        self.spec, self.slices = spec or {}, resource_slices or []
        self.log: List[Dict] = []
        self.taint_ops: List[str] = []
        self.slept = 0
        self._created, self._deleted = set(), set()

    # -- API reads -------------------------------------------------------------------------------
    def driver_enabled(self) -> Tuple[bool, str]:
        if self.spec.get("cluster_policy_error"):
            return False, "failed to get 'cluster-policy': " + self.spec["cluster_policy_error"]
        cp = self.spec.get("cluster_policy")
        if not isinstance(cp, dict):
            return False, ""
        if not isinstance(cp.get("driver_enabled"), bool):
            return False, "'cluster-policy' nvidia container driver configuration (spec.driver.enabled) is not set"
        return cp["driver_enabled"], ""

    def pods(self) -> List[Dict]:
        return [p for p in (self.spec.get("pods") or []) if isinstance(p, dict)]

    def driver_pod(self, node: str):
        labelled = [p for p in self.pods() if (p.get("labels") or {}).get("app.kubernetes.io/component") == "nvidia-driver"]
        if not any(p.get("node") == node for p in labelled):
            return None, "no Pod with label 'app.kubernetes.io/component=nvidia-driver' found on node " + node
        return labelled[0], ""

    def plugin_pod(self, node: str):
        for p in self.pods():
            if (p.get("labels") or {}).get("app.kubernetes.io/name") == "nvidia-dra-driver-gpu" and p.get("node") == node and \
                    p.get("name", "").startswith("nvidia-dra-driver-gpu-kubelet-plugin"):
                return p, ""
        return None, "no Pod named 'nvidia-dra-driver-gpu-kubelet-plugin' found on node " + node

    def agent_pod(self, node: str):
        for p in self.pods():
            if (p.get("labels") or {}).get("app") == "cro-node-agent" and p.get("node") == node and p.get("name", "").startswith("cro-node-agent"):
                return p, ""
        return None, "no Pod named 'cro-node-agent' found on node " + node

    def slice_uuids(self) -> List[str]:
        out = []
        for s in self.slices:
            for d in s.get("devices") or []:
                a = d.get("attributes")
                if isinstance(a, dict) and "uuid" in a:
                    out.append(a["uuid"] if isinstance(a["uuid"], str) else "")
        return out

    # -- pod exec --------------------------------------------------------------------------------
    def run(self, pod: Dict, argv: List[str], kind: str = "command", detached: bool = False) -> Tuple[str, str, Optional[str]]:
        container = (pod.get("containers") or [""])[0]
        query = raw_query(argv, container)
        self.log.append({"pod": pod.get("namespace", "") + "/" + pod.get("name", ""), "container": container, "query": query,
                         "argv": argv, "kind": kind, "detached": detached})
        for rule in self.spec.get("exec") or []:
            if not isinstance(rule, dict):
                continue
            n = rule.get("needle")
            if isinstance(n, dict):
                needle = _escape(n["escape"]) if isinstance(n.get("escape"), str) else n.get("literal", "") if isinstance(n.get("literal"), str) else ""
                if needle not in query:
                    continue
            return rule.get("stdout", ""), rule.get("stderr", ""), rule.get("exec_err") if isinstance(rule.get("exec_err"), str) else None
        return "", "", "no exec rule matches " + query

    def scan(self, pod: Dict, kind: str, prefix: List[str], target: str = "") -> Tuple[str, str, Optional[str]]:
        text = {"fd_scan": 'TARGET_FILE="%s"; <open-file scan of that device node, answered natively>' % target,
                "proc_scan": "<scan of /proc/driver/nvidia/gpus/*/information: minor,uuid,bus per line, answered natively>",
                "cmdline_scan": 'TARGET="%s"; <scan of /proc/*/cmdline for a writer of that path, answered natively>' % target}[kind]
        return self.run(pod, prefix + ["/bin/sh" if prefix else "sh", "-c", text], kind)


def _v(err: Optional[str]) -> str:
    return "<nil>" if err is None else err


def _bad(r) -> bool:
    return r[2] is not None or r[1] != ""


def _smi_infos(c: Cluster, pod: Dict, prefix: List[str], query: str):
    r = c.run(pod, prefix + ["/usr/bin/nvidia-smi", "--query-gpu=" + query, "--format=csv,noheader,nounits"])
    p = _o.parse_gpu_csv(r[0], r[1], r[2], query)
    return (None, p.error) if p.code != 0 else (p.infos or [], "")


def infos_from_driver_pod(c: Cluster, node: str, query: str):
    pod, err = c.driver_pod(node)
    if err:
        return None, err
    return _smi_infos(c, pod, [], query)


def infos_from_proc(c: Cluster, pod: Dict, query: str):
    r = c.scan(pod, "proc_scan", ["/bin/chroot", "/host-root"])
    p = _o.parse_proc_csv(r[0], r[1], r[2], query)
    return (None, p.error) if p.code != 0 else (p.infos or [], "")


def run_nvidia_smi(c: Cluster, node: str) -> str:
    enabled, err = c.driver_enabled()
    if err:
        return err
    if enabled:
        return infos_from_driver_pod(c, node, "gpu_uuid")[1]
    pod, err = c.agent_pod(node)
    if err:
        return err
    return _smi_infos(c, pod, ["/bin/chroot", "/host-root"], "gpu_uuid")[1]


def check_gpu_visible(c: Cluster, device_resource_type: str, node: str, device_id: str) -> Tuple[bool, str]:
    if device_resource_type == "DRA":
        return device_id in c.slice_uuids(), ""
    infos, err = infos_from_driver_pod(c, node, "gpu_uuid")
    if err:
        return False, err
    return any(g.get("gpu_uuid") == device_id for g in infos), ""


def check_no_gpu_loads(c: Cluster, node: str, target_uuid: Optional[str]) -> str:
    enabled, err = c.driver_enabled()
    if err:
        return err
    if not enabled:
        pod, err = c.agent_pod(node)
        if err:
            return err
        infos, err = infos_from_proc(c, pod, "gpu_uuid")
        if err:
            return err
        if target_uuid is None:
            return "runtime error: invalid memory address or nil pointer dereference"
        if not any(g.get("gpu_uuid") == target_uuid for g in infos):
            return ""
        argv = ["/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "--query-compute-apps=gpu_uuid,process_name", "--format=csv,noheader,nounits"]
    else:
        pod, err = c.driver_pod(node)
        if err:
            return ""
        argv = ["/usr/bin/nvidia-smi", "--query-compute-apps=gpu_uuid,process_name", "--format=csv,noheader,nounits"]
    r = c.run(pod, argv)
    return _o.check_no_gpu_loads(r[0], r[1], r[2], pod.get("name", ""), node, target_uuid, enabled)


def _step_failed(verb: str, desc: str, r) -> str:
    return "%s command '%s' failed: '%s', stderr: '%s', stdout: '%s'" % (verb, desc, _v(r[2]), r[1], r[0])


def _remove_modules(c: Cluster, pod: Dict) -> str:
    r = c.run(pod, ["/bin/chroot", "/host-root", "/usr/sbin/lsmod"])
    if r[1] != "" or r[2] is not None:
        return "detach command 'lsmod' failed: '%s', stderr: '%s', stdout: '%s'" % (_v(r[2]), r[1], r[0])
    first = [(line.split() or [""])[0] for line in _o.go_trim_space(r[0]).split("\n")]      # strings.Fields(line)[0]
    for mod in [m for m in ("nvidia_drm", "nvidia_uvm") if m in first]:
        m = c.run(pod, ["/bin/chroot", "/host-root", "/usr/sbin/modprobe", "-r", mod])
        if _bad(m):
            return _step_failed("detach", "remove %s module" % mod, m)
    return ""


def _reset_running(c: Cluster, pod: Dict, sysfs_bus: str) -> Tuple[bool, str]:
    r = c.scan(pod, "cmdline_scan", ["/bin/chroot", "/host-root"], "/sys/bus/pci/devices/%s/remove" % sysfs_bus)
    if r[1] != "" or r[2] is not None:
        return False, "check 'reset GPU' command failed: '%s', stderr: '%s'" % (_v(r[2]), r[1])
    return _o.go_trim_space(r[0]) == "true", ""


def drain_gpu(c: Cluster, node: str, uuid: str, device_resource_type: str) -> str:
    enabled, err = c.driver_enabled()
    if err:
        return err
    chroot = ["/bin/chroot", "/host-root"]
    if device_resource_type == "DRA" and not enabled:
        pod, err = c.agent_pod(node)
        if err:
            return err
        infos, err = infos_from_proc(c, pod, "device_minor,gpu_uuid,pci.bus_id")
        if err:
            return err
        hit = next((g for g in infos if g.get("gpu_uuid") == uuid), None)
        if hit is None:
            return ""
        minor, bus = hit.get("device_minor", ""), _o.go_trim_space(hit.get("pci.bus_id", "")).upper()
        if _o.go_trim_space(bus) == "":
            return "target GPU bus ID is empty"
        r = c.run(pod, chroot + ["/usr/bin/nvidia-smi", "drain", "-p", _o.go_trim_space(bus), "-q"])
        draining, err = _o.check_gpu_drain_status(r[0], r[1], r[2], node, bus)
        if err:
            return err
        if not draining:
            r = c.run(pod, chroot + ["/usr/bin/nvidia-smi", "-i", uuid, "-pm", "0"])
            if _bad(r):
                return _step_failed("deatch", "disable persistence mode", r)
        r = c.scan(pod, "fd_scan", chroot, "/dev/nvidia" + minor)
        if _bad(r):
            return _step_failed("deatch", "check /dev/nvidiaX", r)
        if r[0] != "":
            return "check /dev/nvidiaX command failed: /dev/nvidiaX is in use by one or more processes: " + r[0]
        if not draining:
            r = c.run(pod, chroot + ["/usr/bin/nvidia-smi", "drain", "-p", bus, "-m", "1"])
            if _bad(r):
                return _step_failed("deatch", "set maintenance mode", r)
        r = c.run(pod, chroot + ["/usr/bin/rm", "-f", "/dev/nvidia" + minor])
        if _bad(r):
            return _step_failed("deatch", "remove file /dev/nvidiaX", r)
        if len(infos) != 1:
            r = c.run(pod, chroot + ["/usr/bin/nvidia-smi", "drain", "-p", bus, "-r"])
            return "detach command 'reset GPU' failed: '%s', stderr: '%s', stdout: '%s'" % (_v(r[2]), r[1], r[0]) if _bad(r) else ""
        err = _remove_modules(c, pod)
        if err:
            return err
        sysfs = bus.lower()
        running, err = _reset_running(c, pod, sysfs)
        if err:
            return err
        reset_error = False
        if not running:
            r = c.run(pod, chroot + ["/bin/sh", "-c", "/usr/bin/echo 1 | /usr/bin/tee /sys/bus/pci/devices/%s/remove > /dev/null" % sysfs], detached=True)
            reset_error = _bad(r)
        err = _remove_modules(c, pod)
        if err:
            return err
        c.slept += 1
        running, err = _reset_running(c, pod, sysfs)
        if err:
            return err
        if not running and not reset_error:
            return ""
        return ("detach command 'reset GPU' did not complete, so it failed to drain the last GPU: targetNodeName=%s, targetGPUUUID=%s, "
                "resetCommandRunning=%s, resetCommandError=%s" % (node, uuid, str(running).lower(), str(reset_error).lower()))

    dra = device_resource_type == "DRA"
    driver, err = c.driver_pod(node)
    if err:
        return "" if dra else err
    infos, err = infos_from_driver_pod(c, node, "device_minor,gpu_uuid,pci.bus_id")
    if err:
        return err
    hit = next((g for g in infos if g.get("gpu_uuid") == uuid), None)
    minor = hit.get("device_minor", "") if hit else ""
    bus = hit.get("pci.bus_id", "") if hit else ""
    if bus.startswith("0000"):
        bus = bus[4:]
    if bus == "":
        return ""
    r = c.run(driver, ["/usr/bin/nvidia-smi", "-i", uuid, "-pm", "0"])
    if _bad(r):
        return _step_failed("deatch", "disable persistence mode", r)
    r = c.scan(driver, "fd_scan", [], "/dev/nvidia" + minor)
    err = _o.check_device_file_scan(r[0], r[1], r[2], False)
    if err:
        return err
    if dra:
        r = c.run(driver, ["/usr/bin/rm", "-f", "/run/nvidia/driver/dev/nvidia" + minor])
        if _bad(r):
            return _step_failed("delete device file", "remove file /run/nvidia/driver/dev/nvidiaX", r)
        plugin, err = c.plugin_pod(node)
        if err:
            return err
        r = c.run(plugin, ["/usr/bin/rm", "-f", "/dev/nvidia" + minor])
        if _bad(r):
            return _step_failed("delete device file", "remove file /dev/nvidiaX", r)
    r = c.run(driver, ["/usr/bin/nvidia-smi", "drain", "-p", bus, "-m", "1"])
    if _bad(r):
        return _step_failed("detach", "set maintenance mode", r)
    c.run(driver, ["/usr/bin/nvidia-smi", "drain", "-p", bus, "-r"])
    return ""
