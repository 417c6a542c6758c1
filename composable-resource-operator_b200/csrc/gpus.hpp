// gpus.hpp — the node-side operations of the attach / detach steps
// (internal/utils/gpus.go) over two seams instead of a kubelet:
//
//   Kube  — the API reads gpus.go issues (ClusterPolicy, pod lists)
//   Exec  — "run this in that pod's first container" (execCommandInPod, :788-815)
//
// What is restated: which pod is chosen and what each branch does when it is
// missing, the argv of every nvidia-smi / rm / modprobe / lsmod call, the order
// of the steps of DrainGPU in its three flavours (RKE2+DRA :196-386, OCP+DRA
// :387-548, OCP+DEVICE_PLUGIN :549-664), which failures are fatal, which are
// swallowed, and every error string.  What is NOT copied: the three shell
// scripts the reference ships through `sh -c` (open-file scan of /dev/nvidiaX,
// /proc/driver/nvidia/gpus enumeration, cmdline scan for a running sysfs
// remove).  Those steps are requests of kind FdScan / ProcScan / CmdlineScan;
// libcroprobe answers them natively (detach.cpp, identity.cpp) and a remote
// host may answer them however it likes.  For a scripted Exec (the parity
// harness) they are presented as `sh -c` commands whose text carries the same
// landmark the reference's mocks match on (`TARGET_FILE="/dev/nvidia0"`).
#pragma once

#include <map>
#include <string>
#include <vector>

#include "../../include/croprobe.h"
#include "reconcile.hpp"

namespace cro {
namespace gpus {

using controller::Error;

struct Pod {
    std::string ns, name, node;
    std::map<std::string, std::string> labels;
    std::vector<std::string> containers;     // Spec.Containers[*].Name
};

class Kube {
public:
    virtual ~Kube() {}
    // ClusterPolicy "cluster-policy": *found=false is NotFound; *set=false is spec.driver.enabled == nil
    virtual Error GetClusterPolicy(bool* found, bool* set, bool* enabled) = 0;
    virtual Error ListPods(std::vector<Pod>* out) = 0;                      // every namespace, API order
    virtual Error ListResourceSliceUUIDs(std::vector<std::string>* out) = 0; // attribute "uuid" of every device (DRA)
    // DeviceTaintRule bookkeeping of the DRA detach (gpus.go:691-786); the defaults make taints a no-op
    struct SliceDevice { std::string driver, pool, device, uuid; };
    struct TaintRule { std::string name, driver, pool, device, key, value, effect; };
    virtual Error ListResourceSliceDevices(std::vector<SliceDevice>* /*out*/) { return Error::Nil(); }
    virtual Error GetDeviceTaintRule(const std::string& /*name*/, bool* found) { *found = false; return Error::Nil(); }
    virtual Error CreateDeviceTaintRule(const TaintRule& /*rule*/) { return Error::Nil(); }
    virtual Error DeleteDeviceTaintRule(const std::string& /*name*/) { return Error::Nil(); }
};

struct ExecRequest {
    enum Kind { Command, FdScan, ProcScan, CmdlineScan } kind = Command;
    std::vector<std::string> argv;    // Command: the argv the reference execs.  Scans: the argv prefix (chroot or not)
    std::string target;               // FdScan: "/dev/nvidia0"; CmdlineScan: "/sys/bus/pci/devices/<bus>/remove"
    bool rke2_format = false;         // FdScan: "PID comm, PID comm" (RKE2) vs first holder's comm (OCP)
    bool detached = false;            // fire and forget (the sysfs remove of the last GPU, :334-351)
};
struct ExecResult {
    std::string std_out, std_err;
    bool failed = false;              // execErr != nil
    std::string exec_err;             // its text ("<nil>" is printed when !failed)
};
class Exec {
public:
    virtual ~Exec() {}
    virtual ExecResult Run(const Pod& pod, const std::string& container, const ExecRequest& req) = 0;
    virtual void Sleep(int /*seconds*/) {}   // the 1 s pause before the second cmdline scan (:357)
};

// client-go's exec URL query for a command, as the reference's mock executors see it
// (url.Values.Encode of PodExecOptions: command=..&command=..&container=..&stderr=true&stdout=true).
std::string ExecRawQuery(const std::vector<std::string>& argv, const std::string& container);
// The argv a scan request stands for when it has to look like a command (scripted Exec).
std::vector<std::string> ScanAsCommand(const ExecRequest& req);

class GpuNodeOps : public controller::NodeOps {
public:
    GpuNodeOps(Kube* kube, Exec* exec) : kube_(kube), exec_(exec) {}

    Error CheckNoGPULoads(const std::string& node) override { return CheckNoGPULoadsFor(node, nullptr); }
    Error CheckNoGPULoadsFor(const std::string& node, const std::string* targetGPUUUID) override;   // :88-186
    Error RunNvidiaSmi(const std::string& node) override;                                           // :666-689
    Error CheckGPUVisible(const std::string& deviceResourceType, const controller::ComposableResource& resource,
                          bool* visible) override;                                                  // :54-86
    Error DrainGPU(const std::string& node, const std::string& targetGPUUUID,
                   const std::string& deviceResourceType) override;                                 // :188-664
    Error CreateDeviceTaint(const controller::ComposableResource& resource) override;               // :691-748
    Error DeleteDeviceTaint(const controller::ComposableResource& resource) override;               // :750-766
    // cluster bookkeeping stays with the host
    Error RestartDaemonset(const std::string&, const std::string&) override { return Error::Nil(); }

    // pod choices (:817-876)
    Error getNvidiaDriverDaemonsetPod(const std::string& node, Pod* out);
    Error getDRAKubeletPluginPod(const std::string& node, Pod* out);
    Error getCroNodeAgentPod(const std::string& node, Pod* out);
    Error isContainerDriverEnabled(bool* enabled);                                                  // :1228-1242

protected:
    typedef std::vector<std::map<std::string, std::string>> GpuInfos;
    Error getGPUInfoFromNvidiaPod(const std::string& node, const std::string& query, GpuInfos* out);        // :878-919
    Error getGPUInfoFromCroNodeAgentPod(const std::string& node, const std::string& query, GpuInfos* out);  // :921-962
    Error getGPUInfoFromProc(const Pod& pod, const std::string& query, GpuInfos* out);                      // :1014-1089
    Error checkGPUDrainStatus(const Pod& pod, const std::string& node, const std::string& busID, bool* draining);   // :964-1012
    Error removeNvidiaDriverModule(const Pod& pod);                                                         // :1091-1180
    Error checkResetGPUCommandStillRunning(const Pod& pod, const std::string& busIDForSysfs, bool* running); // :1182-1226
    ExecResult run(const Pod& pod, const ExecRequest& req) { return exec_->Run(pod, pod.containers.empty() ? std::string() : pod.containers[0], req); }

    Kube* kube_;
    Exec* exec_;
};

// ---- the seams answered on the node itself (gpus_local.cpp) --------------------------------------------
// Scans are native /proc walks, `nvidia-smi --query-gpu=...` is answered from already-enumerated devices
// when they are handed in, the compute-apps / drain / persistence-mode invocations go through this process's
// NVML session (nvml_ops.hpp) when there is one, other commands are spawned with the chroot prefix dropped.  Commands that change
// the node (persistence mode, drain -m / -r, rm, modprobe, the sysfs remove) are only run with
// allow_mutation; otherwise they are logged as skipped and succeed, which makes DrainGPU a dry run that
// still performs every read-only check for real.
class LocalExec : public Exec {
public:
    struct Options {
        std::string proc_root;                 // "" = /proc
        bool allow_mutation = false;
        int exec_deadline_ms = 60000;          // per spawned command; expiry = SIGKILL + "context deadline exceeded"
        const cro_dev_info* devs = nullptr;    // devices a probe context enumerated (optional)
        int n_devs = -1;
        bool native_nvml = true;               // answer the detach side's nvidia-smi invocations through NVML (nvml_ops.hpp)
        std::string nvml_lib;                  // "" = libnvidia-ml.so.1
    };
    struct LogEntry { int kind = 0; std::vector<std::string> argv; std::string how; bool failed = false; };
    explicit LocalExec(const Options& o);
    ExecResult Run(const Pod& pod, const std::string& container, const ExecRequest& req) override;
    void Sleep(int seconds) override;
    std::vector<LogEntry> log;

private:
    Options o_;
};

class LocalKube : public Kube {
public:
    LocalKube(const std::string& node, bool driver_container) : node_(node), driver_container_(driver_container) {}
    Error GetClusterPolicy(bool* found, bool* set, bool* enabled) override;
    Error ListPods(std::vector<Pod>* out) override;
    Error ListResourceSliceUUIDs(std::vector<std::string>*) override { return Error::Nil(); }

private:
    std::string node_;
    bool driver_container_;
};

}  // namespace gpus
}  // namespace cro
