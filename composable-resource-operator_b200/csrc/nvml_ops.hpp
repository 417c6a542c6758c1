// nvml_ops.hpp — the nvidia-smi invocations of the detach side, answered through NVML in this process.
//
// The reference spawns nvidia-smi in another pod for each of these (SPDY exec, process start, NVML init — every time):
//   --query-compute-apps=gpu_uuid,process_name --format=csv,noheader,nounits    internal/utils/gpus.go:125,134
//   drain -p <bus> -q                                                           internal/utils/gpus.go:970
//   -i <uuid> -pm 0                                                             internal/utils/gpus.go:267,423,584
//   drain -p <bus> -m 1                                                         internal/utils/gpus.go:269,528,641
//   drain -p <bus> -r                                                           internal/utils/gpus.go:311,529,642
// A node agent that links libcroprobe holds one NVML session and makes the same calls nvidia-smi makes; what comes
// back is shaped like nvidia-smi's stdout / exit code so the reference's unchanged parsers (CheckNoGPULoads :143-167,
// checkGPUDrainStatus :978-1011) consume it.  Only the two queries are pinned against the real nvidia-smi (GPU test
// on the box); the texts of the three mutating commands are never parsed by the reference (it looks at stderr and the
// exec error only) and are marked unpinned.
#pragma once
#include <string>
#include <vector>

namespace cro {
namespace nvml {

struct Reply {
    bool available = false;   // false: no NVML in this process (library or a symbol missing, init failed) — spawn instead
    int exit_code = 0;        // nvidia-smi's: the nvmlReturn_t of the failing call, 0 on success
    std::string std_out;
};

// `lib` = path of libnvidia-ml ("" = libnvidia-ml.so.1); each distinct path keeps its own session for the process.
// --query-gpu=<query> --format=csv,noheader,nounits (gpus.go:886,929) for a caller without a probe context; fields:
// gpu_uuid, device_minor, pci.bus_id, name, index.  Another field: not available (spawn the real nvidia-smi).
Reply QueryGpu(const std::string& lib, const std::string& query);
Reply ComputeApps(const std::string& lib);                                         // gpu_uuid, process_name
Reply DrainQuery(const std::string& lib, const std::string& bus_id);               // drain -p <bus> -q
Reply DrainModify(const std::string& lib, const std::string& bus_id, bool on);     // drain -p <bus> -m 0|1
Reply DrainRemove(const std::string& lib, const std::string& bus_id);              // drain -p <bus> -r
Reply SetPersistence(const std::string& lib, const std::string& gpu, bool on);     // -i <uuid|index|bus> -pm 0|1

// "0000:1F:00.0" / "00000000:1f:00.0" → domain, bus, device (function ignored); false if it is not a PCI address.
bool ParseBusId(const std::string& text, unsigned* domain, unsigned* bus, unsigned* device);

}  // namespace nvml
}  // namespace cro
