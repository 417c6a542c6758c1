// reconcile.hpp — host-side mirror of the reference's attach step, the caller
// of the probe.  Names, argument meaning and error strings follow
//   internal/controller/composableresource_controller.go:176-287,423-441
//   internal/cdi/client.go:25-44            (CdiProvider, sentinel errors)
//   internal/cdi/fti/fm/client.go:184-213   (FM res_op_status gate)
//   api/v1alpha1/composableresource_types.go:27-41
// so the parity tests read like the reference's Ginkgo entries.  (The Go
// toolchain is absent in the build image, hence C++ above the C ABI.)
#pragma once
#include <functional>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "gojson.hpp"

namespace cro {
namespace controller {

// Go `error`: ok() == (err == nil)
struct Error {
    bool set = false;
    std::string msg;
    static Error Nil() { return Error(); }
    static Error New(const std::string& m) { Error e; e.set = true; e.msg = m; return e; }
    bool ok() const { return !set; }
    // A Go run-time panic, restated as the text the runtime would print ("runtime error: index out of range [1] with
    // length 1", ...).  It is NOT an ordinary error: in the reference it unwinds through the handler and through
    // requeueOnErr — no Status().Update is issued on the way — until controller-runtime's Reconcile wrapper recovers it
    // (sigs.k8s.io/controller-runtime v0.21.0 pkg/internal/controller/controller.go: `err = fmt.Errorf("panic: %v
    // [recovered]", r)`; RecoverPanic defaults to true).  Callers test panicked() after every call that may produce one
    // and return Recovered() at once.
    bool panicked() const { return set && msg.compare(0, 15, "runtime error: ") == 0; }
    bool recovered() const { return set && msg.compare(0, 7, "panic: ") == 0; }
    static Error Recovered(const Error& e) { return e.recovered() ? e : New("panic: " + e.msg + " [recovered]"); }
};

extern const std::string ErrWaitingDeviceAttaching;  // "device is attaching to the cluster"
extern const std::string ErrWaitingDeviceDetaching;  // "device is detaching from the cluster"

struct ComposableResourceSpec {
    std::string Type, Model, TargetNode;
    bool ForceDetach = false;
};
struct ComposableResourceStatus {
    std::string State, Error, DeviceID, CDIDeviceID;
    std::string MarshalJSON() const;   // bytes of json.Marshal(status)
};
struct ComposableResource {
    std::string Name;
    std::map<std::string, std::string> Labels;
    bool DeletionTimestampSet = false;
    bool DeleteRequested = false;      // the reconciler called r.Delete(resource) on itself (:298)
    ComposableResourceSpec Spec;
    ComposableResourceStatus Status;
};

struct Result {
    long long RequeueAfterSeconds = 0;
};

// internal/cdi/client.go:34-39
class CdiProvider {
public:
    virtual ~CdiProvider() {}
    virtual Error AddResource(const ComposableResource& instance, std::string* deviceID,
                              std::string* CDIDeviceID) = 0;
    virtual Error RemoveResource(ComposableResource&) { return Error::Nil(); }   // CM records a reason in Status.Error (cm/client.go:201-207)
    virtual Error CheckResource(const ComposableResource&) { return Error::Nil(); }
};

// The FM gate (fti/fm/client.go:184-213) over a ScaleUpResponse body.
Error FMScaleUpResponseToIDs(const std::string& body, const std::string& instanceName,
                             const std::string& specType, const std::string& specModel,
                             std::string* deviceID, std::string* CDIDeviceID);

// CM flavour: checkAddingResources + isSpecMatch + findAvailableDevice
// (internal/cdi/fti/cm/client.go:432-459, 485-509) over the machine JSON of
// GET .../machines/<id> (internal/cdi/fti/cm/api/machine.go:19-93).
// existingDeviceIDs = Status.DeviceID of every ComposableResource in the cluster.
struct CMAddingResult {
    std::string specUUID;
    long long deviceCount = 0;
    std::string deviceID, CDIDeviceID;
    Error err;
};
CMAddingResult CMCheckAddingResources(const std::string& machineBody,
                                      const std::vector<std::string>& existingDeviceIDs,
                                      const std::string& specType, const std::string& specModel);

// Node-side operations the attach step calls (internal/utils): the CUDA probe
// backs RunNvidiaSmi / CheckGPUVisible; the daemonset restarts and the load
// check are cluster bookkeeping, injected.
class NodeOps {
public:
    virtual ~NodeOps() {}
    virtual Error CheckNoGPULoads(const std::string& node) = 0;
    virtual Error RestartDaemonset(const std::string& ns, const std::string& name) = 0;
    virtual Error RunNvidiaSmi(const std::string& node) = 0;
    virtual Error CheckGPUVisible(const std::string& deviceResourceType,
                                  const ComposableResource& resource, bool* visible) = 0;
    // detach side (internal/utils/gpus.go:88, 188, 691, 756); targetGPUUUID null = whole node
    virtual Error CheckNoGPULoadsFor(const std::string& /*node*/, const std::string* /*targetGPUUUID*/) { return Error::Nil(); }
    virtual Error CreateDeviceTaint(const ComposableResource&) { return Error::Nil(); }
    virtual Error DeleteDeviceTaint(const ComposableResource&) { return Error::Nil(); }
    virtual Error DrainGPU(const std::string& /*node*/, const std::string& /*deviceID*/,
                           const std::string& /*deviceResourceType*/) { return Error::Nil(); }
};

class ComposableResourceReconciler {
public:
    ComposableResourceReconciler(CdiProvider* provider, NodeOps* node) : provider_(provider), node_(node) {}
    // composableresource_controller.go:176-198
    Error handleNoneState(ComposableResource* resource, Result* result);
    // composableresource_controller.go:200-287
    Error handleAttachingState(ComposableResource* resource, const std::string& deviceResourceType,
                               Result* result);
    // composableresource_controller.go:289-318
    Error handleOnlineState(ComposableResource* resource, Result* result);
    // composableresource_controller.go:320-407
    Error handleDetachingState(ComposableResource* resource, const std::string& deviceResourceType,
                               Result* result);
    // Every Status().Update the reference would issue, in order (attempts: a failing one is listed too).
    std::vector<ComposableResourceStatus> statusUpdates;
    int failedUpdates = 0;
    // The API server's answer to Status().Update; unset = every write succeeds.  The handlers stop where the
    // reference stops when a write fails (`if err := r.Status().Update(...); err != nil { return r.requeueOnErr(...) }`,
    // composableresource_controller.go:233-235, :248-250, ...).
    std::function<Error(const ComposableResource&)> writer;
    // composableresource_controller.go:423-433
    Error requeueOnErr(ComposableResource* resource, const Error& err);

private:
    Error statusUpdate(const ComposableResource& r) {
        statusUpdates.push_back(r.Status);
        Error e = writer ? writer(r) : Error::Nil();
        if (!e.ok()) ++failedUpdates;
        return e;
    }
    CdiProvider* provider_;
    NodeOps* node_;
};

}  // namespace controller
}  // namespace cro
