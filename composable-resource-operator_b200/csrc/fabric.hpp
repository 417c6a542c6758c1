// fabric.hpp — the fabric manager's wire formats as a reusable codec
// (SURVEY.md §8f rank 3): response-side decode + the decisions the providers
// take on it.  The HTTP transport, OAuth2 and the node->machine lookups are out
// of scope (external appliance); bodies come in as text.
//   FM  CheckResource  internal/cdi/fti/fm/client.go:314-359
//   FM  GetResources   internal/cdi/fti/fm/client.go:361-414   (decode of one node's machine)
//   CM  CheckResource  internal/cdi/fti/cm/client.go:262-304
//   CM  GetResources   internal/cdi/fti/cm/client.go:306-346
// Wire structs: internal/cdi/fti/fm/api/get.go:19-47, internal/cdi/fti/cm/api/machine.go:19-93,
// cdi.DeviceInfo internal/cdi/client.go:25-32.
#pragma once
#include <string>
#include <vector>

#include "reconcile.hpp"

namespace cro {
namespace fabric {

using controller::Error;

struct DeviceInfo {   // internal/cdi/client.go:25-32
    std::string NodeName, MachineUUID, DeviceType, Model, DeviceID, CDIDeviceID;
};

// body = the JSON of GET fabric_manager/api/v1/machines/<id> (GetMachineResponse).
Error FMCheckResource(const std::string& body, const std::string& specType, const std::string& specModel,
                      const std::string& deviceID);
Error FMGetResources(const std::string& body, const std::string& nodeName, const std::string& machineID,
                     std::vector<DeviceInfo>* out);
// body = the JSON of GET cluster_manager/.../machines/<id> (MachineData).
Error CMCheckResource(const std::string& body, const std::string& specType, const std::string& specModel,
                      const std::string& deviceID);
Error CMGetResources(const std::string& body, const std::string& nodeName, const std::string& machineID,
                     std::vector<DeviceInfo>* out);

// The list as the UpstreamSyncer tick consumes it (cro_sim_sync_upstream).
std::string DeviceInfosToJson(const std::vector<DeviceInfo>& v);

}  // namespace fabric
}  // namespace cro
