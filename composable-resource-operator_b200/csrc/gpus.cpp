// gpus.cpp — see gpus.hpp.
#include "gpus.hpp"

#include <cctype>

#include "detach.hpp"
#include "identity.hpp"

namespace cro {
namespace gpus {

namespace {

const char* cstr_or_null(const ExecResult& r) { return r.failed ? r.exec_err.c_str() : nullptr; }
std::string errText(const ExecResult& r) { return r.failed ? r.exec_err : std::string("<nil>"); }   // fmt %v of a nil error
bool bad(const ExecResult& r) { return r.failed || !r.std_err.empty(); }

std::string queryEscape(const std::string& s) {   // net/url.QueryEscape
    static const char hex[] = "0123456789ABCDEF";
    std::string o;
    for (unsigned char c : s) {
        if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '-' || c == '_' || c == '.' || c == '~')
            o.push_back((char)c);
        else if (c == ' ')
            o.push_back('+');
        else { o.push_back('%'); o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
    }
    return o;
}

bool hasPrefix(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }

ExecRequest command(std::vector<std::string> argv) {
    ExecRequest r;
    r.argv = std::move(argv);
    return r;
}

}  // namespace

std::string ExecRawQuery(const std::vector<std::string>& argv, const std::string& container) {
    std::string q;
    for (const auto& a : argv) {
        if (!q.empty()) q += "&";
        q += "command=" + queryEscape(a);
    }
    q += std::string(q.empty() ? "" : "&") + "container=" + queryEscape(container) + "&stderr=true&stdout=true";
    return q;
}

std::vector<std::string> ScanAsCommand(const ExecRequest& req) {
    std::vector<std::string> argv = req.argv;                      // chroot prefix or nothing
    argv.push_back(argv.empty() ? "sh" : "/bin/sh");
    argv.push_back("-c");
    switch (req.kind) {
        case ExecRequest::FdScan:
            argv.push_back("TARGET_FILE=\"" + req.target + "\"; <open-file scan of that device node, answered natively>");
            break;
        case ExecRequest::ProcScan:
            argv.push_back("<scan of /proc/driver/nvidia/gpus/*/information: minor,uuid,bus per line, answered natively>");
            break;
        case ExecRequest::CmdlineScan:
            argv.push_back("TARGET=\"" + req.target + "\"; <scan of /proc/*/cmdline for a writer of that path, answered natively>");
            break;
        default: break;
    }
    return argv;
}

// ---------------------------------------------------------------------------
// API reads
// ---------------------------------------------------------------------------
Error GpuNodeOps::isContainerDriverEnabled(bool* enabled) {
    bool found = false, set = false, en = false;
    Error e = kube_->GetClusterPolicy(&found, &set, &en);
    if (!e.ok()) return Error::New("failed to get 'cluster-policy': " + e.msg);
    if (!found) { *enabled = false; return Error::Nil(); }        // NotFound: the container driver is disabled (RKE2)
    if (!set) return Error::New("'cluster-policy' nvidia container driver configuration (spec.driver.enabled) is not set");
    *enabled = en;
    return Error::Nil();
}

Error GpuNodeOps::getNvidiaDriverDaemonsetPod(const std::string& node, Pod* out) {
    std::vector<Pod> pods;
    Error e = kube_->ListPods(&pods);
    if (!e.ok()) return Error::New("failed to list pods: " + e.msg);
    std::vector<const Pod*> labelled;
    for (const auto& p : pods) {
        auto it = p.labels.find("app.kubernetes.io/component");
        if (it != p.labels.end() && it->second == "nvidia-driver") labelled.push_back(&p);
    }
    bool onNode = false;
    for (const Pod* p : labelled)
        if (p->node == node) onNode = true;
    if (!onNode) return Error::New("no Pod with label 'app.kubernetes.io/component=nvidia-driver' found on node " + node);
    *out = *labelled[0];      // the reference returns Items[0], not the pod it matched (:838)
    return Error::Nil();
}

Error GpuNodeOps::getDRAKubeletPluginPod(const std::string& node, Pod* out) {
    std::vector<Pod> pods;
    Error e = kube_->ListPods(&pods);
    if (!e.ok()) return Error::New("failed to list pods: " + e.msg);
    for (const auto& p : pods) {
        auto it = p.labels.find("app.kubernetes.io/name");
        if (it == p.labels.end() || it->second != "nvidia-dra-driver-gpu" || p.node != node) continue;
        if (hasPrefix(p.name, "nvidia-dra-driver-gpu-kubelet-plugin")) { *out = p; return Error::Nil(); }
    }
    return Error::New("no Pod named 'nvidia-dra-driver-gpu-kubelet-plugin' found on node " + node);
}

Error GpuNodeOps::getCroNodeAgentPod(const std::string& node, Pod* out) {
    std::vector<Pod> pods;
    Error e = kube_->ListPods(&pods);
    if (!e.ok()) return Error::New("failed to list pods: " + e.msg);
    for (const auto& p : pods) {
        auto it = p.labels.find("app");
        if (it == p.labels.end() || it->second != "cro-node-agent" || p.node != node) continue;
        if (hasPrefix(p.name, "cro-node-agent")) { *out = p; return Error::Nil(); }
    }
    return Error::New("no Pod named 'cro-node-agent' found on node " + node);
}

// ---------------------------------------------------------------------------
// enumeration
// ---------------------------------------------------------------------------
Error GpuNodeOps::getGPUInfoFromNvidiaPod(const std::string& node, const std::string& query, GpuInfos* out) {
    Pod pod;
    Error e = getNvidiaDriverDaemonsetPod(node, &pod);
    if (!e.ok()) return e;
    ExecResult r = run(pod, command({"/usr/bin/nvidia-smi", "--query-gpu=" + query, "--format=csv,noheader,nounits"}));
    identity::GpuInfoResult g = identity::getGPUInfoFromNvidiaSmiOutput(r.std_out, r.std_err, cstr_or_null(r), query);
    if (g.code != 0) return Error::New(g.error);
    *out = g.infos;
    return Error::Nil();
}

Error GpuNodeOps::getGPUInfoFromCroNodeAgentPod(const std::string& node, const std::string& query, GpuInfos* out) {
    Pod pod;
    Error e = getCroNodeAgentPod(node, &pod);
    if (!e.ok()) return e;
    ExecResult r = run(pod, command({"/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "--query-gpu=" + query,
                                     "--format=csv,noheader,nounits"}));
    identity::GpuInfoResult g = identity::getGPUInfoFromNvidiaSmiOutput(r.std_out, r.std_err, cstr_or_null(r), query);
    if (g.code != 0) return Error::New(g.error);
    *out = g.infos;
    return Error::Nil();
}

Error GpuNodeOps::getGPUInfoFromProc(const Pod& pod, const std::string& query, GpuInfos* out) {
    ExecRequest req;
    req.kind = ExecRequest::ProcScan;
    req.argv = {"/bin/chroot", "/host-root"};
    ExecResult r = run(pod, req);
    identity::GpuInfoResult g = identity::getGPUInfoFromProcOutput(r.std_out, r.std_err, cstr_or_null(r), query);
    if (g.code != 0) return Error::New(g.error);
    *out = g.infos;
    return Error::Nil();
}

Error GpuNodeOps::RunNvidiaSmi(const std::string& node) {
    bool driverEnabled = false;
    Error e = isContainerDriverEnabled(&driverEnabled);
    if (!e.ok()) return e;
    GpuInfos ignored;
    return driverEnabled ? getGPUInfoFromNvidiaPod(node, "gpu_uuid", &ignored) : getGPUInfoFromCroNodeAgentPod(node, "gpu_uuid", &ignored);
}

Error GpuNodeOps::CheckGPUVisible(const std::string& deviceResourceType, const controller::ComposableResource& resource,
                                  bool* visible) {
    *visible = false;
    if (deviceResourceType == "DRA") {
        std::vector<std::string> uuids;
        Error e = kube_->ListResourceSliceUUIDs(&uuids);
        if (!e.ok()) return e;
        for (const auto& u : uuids)
            if (u == resource.Status.DeviceID) { *visible = true; break; }
        return Error::Nil();
    }
    GpuInfos infos;
    Error e = getGPUInfoFromNvidiaPod(resource.Spec.TargetNode, "gpu_uuid", &infos);
    if (!e.ok()) return e;
    for (auto& g : infos)
        if (g["gpu_uuid"] == resource.Status.DeviceID) { *visible = true; break; }
    return Error::Nil();
}

// ---------------------------------------------------------------------------
// load check
// ---------------------------------------------------------------------------
Error GpuNodeOps::CheckNoGPULoadsFor(const std::string& node, const std::string* targetGPUUUID) {
    bool driverEnabled = false;
    Error e = isContainerDriverEnabled(&driverEnabled);
    if (!e.ok()) return e;
    Pod pod;
    std::vector<std::string> argv;
    if (!driverEnabled) {
        e = getCroNodeAgentPod(node, &pod);
        if (!e.ok()) return e;
        GpuInfos infos;
        e = getGPUInfoFromProc(pod, "gpu_uuid", &infos);
        if (!e.ok()) return e;
        if (!targetGPUUUID)   // the reference dereferences a nil *string here (:115)
            return Error::New("runtime error: invalid memory address or nil pointer dereference");
        bool found = false;
        for (auto& g : infos)
            if (g["gpu_uuid"] == *targetGPUUUID) { found = true; break; }
        if (!found) return Error::Nil();          // already reset: no load by definition
        argv = {"/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "--query-compute-apps=gpu_uuid,process_name", "--format=csv,noheader,nounits"};
    } else {
        if (!getNvidiaDriverDaemonsetPod(node, &pod).ok()) return Error::Nil();   // no driver pod: no GPU, no load (:136-139)
        argv = {"/usr/bin/nvidia-smi", "--query-compute-apps=gpu_uuid,process_name", "--format=csv,noheader,nounits"};
    }
    ExecResult r = run(pod, command(argv));
    return detach::CheckNoGPULoadsFromOutput(r.std_out, r.std_err, cstr_or_null(r), pod.name, node, targetGPUUUID, driverEnabled);
}

// ---------------------------------------------------------------------------
// drain
// ---------------------------------------------------------------------------
Error GpuNodeOps::checkGPUDrainStatus(const Pod& pod, const std::string& node, const std::string& busID, bool* draining) {
    const std::string b = identity::TrimSpace(busID);
    if (b.empty()) return Error::New("target GPU bus ID is empty");
    ExecResult r = run(pod, command({"/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "drain", "-p", b, "-q"}));
    return detach::checkGPUDrainStatusFromOutput(r.std_out, r.std_err, cstr_or_null(r), node, busID, draining);
}

Error GpuNodeOps::removeNvidiaDriverModule(const Pod& pod) {
    ExecResult r = run(pod, command({"/bin/chroot", "/host-root", "/usr/sbin/lsmod"}));
    if (!r.std_err.empty() || r.failed)
        return Error::New("detach command 'lsmod' failed: '" + errText(r) + "', stderr: '" + r.std_err + "', stdout: '" + r.std_out + "'");
    bool drm = false, uvm = false;
    for (const auto& line : identity::Split(identity::TrimSpace(r.std_out), "\n")) {
        // strings.Fields(line)[0]
        size_t i = 0;
        while (i < line.size() && isspace((unsigned char)line[i])) ++i;
        size_t j = i;
        while (j < line.size() && !isspace((unsigned char)line[j])) ++j;
        const std::string first = line.substr(i, j - i);
        if (first == "nvidia_drm") drm = true;
        if (first == "nvidia_uvm") uvm = true;
    }
    std::vector<std::pair<std::string, std::string>> steps;   // (module, description)
    if (drm) steps.push_back({"nvidia_drm", "remove nvidia_drm module"});
    if (uvm) steps.push_back({"nvidia_uvm", "remove nvidia_uvm module"});
    for (const auto& s : steps) {
        ExecResult m = run(pod, command({"/bin/chroot", "/host-root", "/usr/sbin/modprobe", "-r", s.first}));
        if (bad(m))
            return Error::New("detach command '" + s.second + "' failed: '" + errText(m) + "', stderr: '" + m.std_err + "', stdout: '" + m.std_out + "'");
    }
    return Error::Nil();
}

Error GpuNodeOps::checkResetGPUCommandStillRunning(const Pod& pod, const std::string& busIDForSysfs, bool* running) {
    ExecRequest req;
    req.kind = ExecRequest::CmdlineScan;
    req.argv = {"/bin/chroot", "/host-root"};
    req.target = "/sys/bus/pci/devices/" + busIDForSysfs + "/remove";
    ExecResult r = run(pod, req);
    if (!r.std_err.empty() || r.failed)
        return Error::New("check 'reset GPU' command failed: '" + errText(r) + "', stderr: '" + r.std_err + "'");
    *running = identity::TrimSpace(r.std_out) == "true";
    return Error::Nil();
}

Error GpuNodeOps::DrainGPU(const std::string& node, const std::string& targetGPUUUID, const std::string& deviceResourceType) {
    bool driverEnabled = false;
    Error e = isContainerDriverEnabled(&driverEnabled);
    if (!e.ok()) return e;
    auto stepFailed = [](const char* verb, const std::string& desc, const ExecResult& r) {
        return Error::New(std::string(verb) + " command '" + desc + "' failed: '" + errText(r) + "', stderr: '" + r.std_err + "', stdout: '" + r.std_out + "'");
    };

    if (deviceResourceType == "DRA" && !driverEnabled) {
        // ---- RKE2 + DRA (:196-386): everything through the cro-node-agent pod, chrooted into the host ----
        Pod agent;
        e = getCroNodeAgentPod(node, &agent);
        if (!e.ok()) return e;
        GpuInfos infos;
        e = getGPUInfoFromProc(agent, "device_minor,gpu_uuid,pci.bus_id", &infos);
        if (!e.ok()) return e;
        bool found = false;
        std::string minor, bus;
        for (auto& g : infos)
            if (g["gpu_uuid"] == targetGPUUUID) {
                minor = g["device_minor"];
                bus = identity::ToUpper(identity::TrimSpace(g["pci.bus_id"]));
                found = true;
                break;
            }
        if (!found) return Error::Nil();         // already reset
        bool draining = false;
        e = checkGPUDrainStatus(agent, node, bus, &draining);
        if (!e.ok()) return e;
        // "disable persistence mode" and "set maintenance mode" are skipped when the GPU is already draining
        if (!draining) {
            ExecResult r = run(agent, command({"/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "-i", targetGPUUUID, "-pm", "0"}));
            if (bad(r)) return stepFailed("deatch", "disable persistence mode", r);
        }
        {
            ExecRequest scan;
            scan.kind = ExecRequest::FdScan;
            scan.argv = {"/bin/chroot", "/host-root"};
            scan.target = "/dev/nvidia" + minor;
            scan.rke2_format = true;
            ExecResult r = run(agent, scan);
            if (bad(r)) return stepFailed("deatch", "check /dev/nvidiaX", r);
            if (!r.std_out.empty())
                return Error::New("check /dev/nvidiaX command failed: /dev/nvidiaX is in use by one or more processes: " + r.std_out);
        }
        if (!draining) {
            ExecResult r = run(agent, command({"/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "drain", "-p", bus, "-m", "1"}));
            if (bad(r)) return stepFailed("deatch", "set maintenance mode", r);
        }
        {
            ExecResult r = run(agent, command({"/bin/chroot", "/host-root", "/usr/bin/rm", "-f", "/dev/nvidia" + minor}));
            if (bad(r)) return stepFailed("deatch", "remove file /dev/nvidiaX", r);
        }
        if (infos.size() != 1) {                  // other GPUs remain: a plain reset
            ExecResult r = run(agent, command({"/bin/chroot", "/host-root", "/usr/bin/nvidia-smi", "drain", "-p", bus, "-r"}));
            if (bad(r)) return Error::New("detach command 'reset GPU' failed: '" + errText(r) + "', stderr: '" + r.std_err + "', stdout: '" + r.std_out + "'");
            return Error::Nil();
        }
        // the last GPU: unload the dependent modules, remove the PCI function through sysfs (fire and
        // forget), unload again, give it a second, and see whether the writer is still there (:300-385)
        e = removeNvidiaDriverModule(agent);
        if (!e.ok()) return e;
        const std::string sysfsBus = identity::ToLower(bus);
        bool running = false;
        e = checkResetGPUCommandStillRunning(agent, sysfsBus, &running);
        if (!e.ok()) return e;
        bool resetError = false;
        if (!running) {
            ExecRequest rm = command({"/bin/chroot", "/host-root", "/bin/sh", "-c",
                                      "/usr/bin/echo 1 | /usr/bin/tee /sys/bus/pci/devices/" + sysfsBus + "/remove > /dev/null"});
            rm.detached = true;
            ExecResult r = run(agent, rm);
            if (bad(r)) resetError = true;        // only logged by the reference's goroutine; decides the verdict below
        }
        e = removeNvidiaDriverModule(agent);
        if (!e.ok()) return e;
        exec_->Sleep(1);
        e = checkResetGPUCommandStillRunning(agent, sysfsBus, &running);
        if (!e.ok()) return e;
        if (!running && !resetError) return Error::Nil();
        return Error::New("detach command 'reset GPU' did not complete, so it failed to drain the last GPU: targetNodeName=" + node +
                          ", targetGPUUUID=" + targetGPUUUID + ", resetCommandRunning=" + (running ? "true" : "false") +
                          ", resetCommandError=" + (resetError ? "true" : "false"));
    }

    // ---- OCP: the nvidia-driver-daemonset pod does the work -------------------------------------------
    const bool dra = deviceResourceType == "DRA";
    Pod driver;
    e = getNvidiaDriverDaemonsetPod(node, &driver);
    if (!e.ok()) return dra ? Error::Nil() : e;   // DRA: no driver pod = no GPU on the node, nothing to drain (:389-393)
    GpuInfos infos;
    e = getGPUInfoFromNvidiaPod(node, "device_minor,gpu_uuid,pci.bus_id", &infos);
    if (!e.ok()) return e;
    std::string minor, bus;
    for (auto& g : infos)
        if (g["gpu_uuid"] == targetGPUUUID) {
            minor = g["device_minor"];
            bus = identity::TrimPrefix(g["pci.bus_id"], "0000");
            break;
        }
    if (bus.empty()) return Error::Nil();         // not enumerated any more: already drained
    {
        ExecResult r = run(driver, command({"/usr/bin/nvidia-smi", "-i", targetGPUUUID, "-pm", "0"}));
        if (bad(r)) return stepFailed("deatch", "disable persistence mode", r);
    }
    {
        ExecRequest scan;
        scan.kind = ExecRequest::FdScan;
        scan.target = "/dev/nvidia" + minor;
        ExecResult r = run(driver, scan);
        Error se = detach::CheckDeviceFileScanResult(r.std_out, r.std_err, cstr_or_null(r), false);
        if (!se.ok()) return se;
    }
    if (dra) {
        // the driver does not delete the device nodes itself: once in the driver's /run tree, once in the
        // kubelet plugin's /dev (:476-520)
        ExecResult r = run(driver, command({"/usr/bin/rm", "-f", "/run/nvidia/driver/dev/nvidia" + minor}));
        if (bad(r)) return stepFailed("delete device file", "remove file /run/nvidia/driver/dev/nvidiaX", r);
        Pod plugin;
        e = getDRAKubeletPluginPod(node, &plugin);
        if (!e.ok()) return e;
        ExecResult p = run(plugin, command({"/usr/bin/rm", "-f", "/dev/nvidia" + minor}));
        if (bad(p)) return stepFailed("delete device file", "remove file /dev/nvidiaX", p);
    }
    {
        ExecResult r = run(driver, command({"/usr/bin/nvidia-smi", "drain", "-p", bus, "-m", "1"}));
        if (bad(r)) return stepFailed("detach", "set maintenance mode", r);
        run(driver, command({"/usr/bin/nvidia-smi", "drain", "-p", bus, "-r"}));   // a failed reset is ignored (:540-542, :656-658)
    }
    return Error::Nil();
}

// ---------------------------------------------------------------------------
// device taints (DRA): keep the scheduler away from a device that is being detached
// ---------------------------------------------------------------------------
Error GpuNodeOps::CreateDeviceTaint(const controller::ComposableResource& resource) {
    const std::string taintName = resource.Name + "-taint";
    bool exists = false;
    Error e = kube_->GetDeviceTaintRule(taintName, &exists);
    if (!e.ok()) return e;                       // any error but NotFound is returned as is (:697-699)
    if (exists) return Error::Nil();
    std::vector<Kube::SliceDevice> devices;
    e = kube_->ListResourceSliceDevices(&devices);
    if (!e.ok()) return e;
    const Kube::SliceDevice* hit = nullptr;
    for (const auto& d : devices)
        if (d.uuid == resource.Status.DeviceID) { hit = &d; break; }
    if (!hit || hit->device.empty()) return Error::Nil();   // the slice no longer lists it: nothing to taint
    Kube::TaintRule rule{taintName, hit->driver, hit->pool, hit->device, "k8s.io/device-uuid", resource.Status.DeviceID, "NoSchedule"};
    e = kube_->CreateDeviceTaintRule(rule);
    if (!e.ok()) return Error::New("failed to create DeviceTaintRule " + taintName + ": " + e.msg);
    return Error::Nil();
}

Error GpuNodeOps::DeleteDeviceTaint(const controller::ComposableResource& resource) {
    const std::string taintName = resource.Name + "-taint";
    bool exists = false;
    Error e = kube_->GetDeviceTaintRule(taintName, &exists);
    if (!e.ok()) return Error::New("failed to get DeviceTaintRule " + taintName + ": " + e.msg);
    if (!exists) return Error::Nil();
    e = kube_->DeleteDeviceTaintRule(taintName);
    if (!e.ok()) return Error::New("failed to delete DeviceTaintRule " + taintName + ": " + e.msg);
    return Error::Nil();
}

}  // namespace gpus
}  // namespace cro
