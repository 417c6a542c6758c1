// detach.hpp — detach-side pre-flight, the step on the other side of the hot
// path.  The reference does these with 3–8 SPDY execs of
// shell / nvidia-smi per detach (internal/utils/gpus.go:88-186, 236-260,
// 441-473, 964-1012); here the text rules are restated quirk-for-quirk and the
// open-file scan is done natively on /proc.
#pragma once
#include <string>

#include "reconcile.hpp"

namespace cro {
namespace detach {

using controller::Error;

// The parse + decision half of utils.CheckNoGPULoads (gpus.go:145-186) over the
// output of `nvidia-smi --query-compute-apps=gpu_uuid,process_name
// --format=csv,noheader,nounits`.  driverEnabled selects the OCP branch
// (any load on the node is an error) vs the RKE2 branch (only loads on
// *targetGPUUUID).  exec_err == nullptr is a nil error.
Error CheckNoGPULoadsFromOutput(const std::string& stdOut, const std::string& stdErr, const char* exec_err,
                                const std::string& podName, const std::string& targetNodeName,
                                const std::string* targetGPUUUID, bool driverEnabled);

// checkGPUDrainStatus (gpus.go:964-1012) over the output of
// `nvidia-smi drain -p <bus> -q`.
Error checkGPUDrainStatusFromOutput(const std::string& stdOut, const std::string& stdErr, const char* exec_err,
                                    const std::string& targetNodeName, const std::string& targetGPUBusID,
                                    bool* draining);

// The decision after the fd-scan script (gpus.go:468-473 / :629-634, OCP
// flavour; :286-291 RKE2 flavour when rke2 is true).
Error CheckDeviceFileScanResult(const std::string& stdOut, const std::string& stdErr, const char* exec_err, bool rke2);

// Native replacement of the fd-scan shell scripts: which processes hold
// `target` (e.g. "/dev/nvidia0") open.  rke2 == false prints what the OCP
// script prints (comm of the first holder + "\n"); rke2 == true prints
// "PID comm, PID comm".  proc_root lets tests point at a fake /proc.
std::string ScanDeviceFileHolders(const std::string& proc_root, const std::string& target, bool rke2);

// Native replacement of the cmdline scan of checkResetGPUCommandStillRunning (gpus.go:1182-1226): is some
// OTHER process's command line mentioning `target` (the sysfs "remove" file a detached `tee` writes to)?
// Prints what the script prints: "true\n" or nothing.
std::string ScanCmdlineFor(const std::string& proc_root, const std::string& target);

}  // namespace detach
}  // namespace cro
