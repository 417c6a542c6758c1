// cluster.hpp — in-memory API server + the two reconcilers, the CALLER of the
// hot path (drives BASELINE configs 4 "reconcile storm"
// and 5 "attach/detach churn").
//
// Restates, with the reference's names and error strings:
//   ComposabilityRequestReconciler   internal/controller/composabilityrequest_controller.go:72-625
//   ComposableResourceReconciler     internal/controller/composableresource_controller.go:73-441
//   CheckNodeCapacitySufficient      internal/utils/nodes.go:78-117 (incl. the milli-CPU vs whole-core quirk)
//   GenerateComposableResourceName   internal/utils/stringutils.go:26-33
// against an in-memory store with Kubernetes delete semantics (finalizers,
// deletionTimestamp) and controller-runtime style de-duplicating work queues.
// Differences from a live cluster, all stated in the stats it returns:
//   * RequeueAfter timers (30 s / 3 s) fire immediately once the queue drains,
//     and stop when a full timer round changes nothing;
//   * attach / detach is logical (CUDA cannot hot-plug inside one process):
//     the fake fabric hands out the UUID of the physical GPU mapped to the node;
//   * map iteration that Go leaves unordered is done in key order.
#pragma once
#include <chrono>
#include <deque>
#include <map>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "gojson.hpp"
#include "reconcile.hpp"

struct cro_ctx;

namespace cro {
namespace sim {

using controller::Error;

struct Node {
    std::string Name;
    long long CPU = 0, Memory = 0, EphemeralStorage = 0, Pods = 0;   // Capacity, Quantity.AsInt64()
};

struct NodeSpec {   // api/v1alpha1/composabilityrequest_types.go:55-64
    long long MilliCPU = 0, Memory = 0, EphemeralStorage = 0, AllowedPodNumber = 0;
    bool operator==(const NodeSpec& o) const {
        return MilliCPU == o.MilliCPU && Memory == o.Memory && EphemeralStorage == o.EphemeralStorage &&
               AllowedPodNumber == o.AllowedPodNumber;
    }
};

struct ScalarResourceDetails {   // composabilityrequest_types.go:40-53
    std::string Type, Model;
    long long Size = 0;
    bool ForceDetach = false;
    std::string AllocationPolicy, TargetNode;
    bool HasOtherSpec = false;
    NodeSpec OtherSpec;
    bool operator==(const ScalarResourceDetails& o) const {
        return Type == o.Type && Model == o.Model && Size == o.Size && ForceDetach == o.ForceDetach &&
               AllocationPolicy == o.AllocationPolicy && TargetNode == o.TargetNode &&
               HasOtherSpec == o.HasOtherSpec && (!HasOtherSpec || OtherSpec == o.OtherSpec);
    }
    std::string MarshalJSON() const;
};

struct ScalarResourceStatus {   // composabilityrequest_types.go:74-80
    std::string State, DeviceID, CDIDeviceID, NodeName, Error;
    bool operator==(const ScalarResourceStatus& o) const {
        return State == o.State && DeviceID == o.DeviceID && CDIDeviceID == o.CDIDeviceID &&
               NodeName == o.NodeName && Error == o.Error;
    }
    std::string MarshalJSON() const;
};

struct ComposabilityRequestStatus {   // composabilityrequest_types.go:66-72
    std::string State, Error;
    std::map<std::string, ScalarResourceStatus> Resources;
    ScalarResourceDetails ScalarResource;
    bool operator==(const ComposabilityRequestStatus& o) const {
        return State == o.State && Error == o.Error && Resources == o.Resources && ScalarResource == o.ScalarResource;
    }
    std::string MarshalJSON() const;
};

struct ComposabilityRequest {
    std::string Name;
    std::vector<std::string> Finalizers;
    bool DeletionTimestampSet = false;
    long long CreationSeq = 0;
    ScalarResourceDetails Spec;   // Spec.Resource
    ComposabilityRequestStatus Status;
};

struct ApiFault {   // thrown by a write the armed fault refuses; caught at the reconcile boundary
    std::string msg;
};

struct StoredResource {
    controller::ComposableResource obj;
    std::map<std::string, std::string> Annotations;
    std::vector<std::string> Finalizers;
    long long CreationSeq = 0;
};

struct Stats {
    long long request_reconciles = 0, resource_reconciles = 0, status_updates = 0, spec_bytes = 0;
    long long probes = 0, probe_failures = 0, timer_rounds = 0, reconcile_errors = 0;
    std::vector<long long> reconcile_ns;   // one entry per Reconcile call
    double wall_s = 0;
    struct Gpu {
        long long probes = 0, begins_behind_running = 0, gaps_over_100us = 0;
        unsigned long long busy_ns = 0, first_start_ns = 0, last_end_ns = 0, gap_ns = 0, max_gap_ns = 0;
    };
    Gpu gpu[16];                           // per probe-context device: how busy the worker kept it (device timers)
    long long blocked_ns = 0;              // worker time spent waiting for a probe with nothing else to reconcile
    struct Ev { long long t_us; char what; int dev; };
    std::vector<Ev> trace;                 // first events of the probe slot: b(egin) c(ollect) w(ait queued) r(elease wakes)
    bool tracing = false;
};

class Cluster {
public:
    // config: {"nodes":[{"name":..,"cpu":..,"memory":..,"ephemeral_storage":..,"pods":..} | "name", ...],
    //          "device_resource_type":"DEVICE_PLUGIN"|"DRA", "probe":bool, "seed":N,
    //          "uuids":[...]  (device of node i = uuids[i mod n]; defaults to the probe context's devices)}
    Cluster(cro_ctx* ctx, const gojson::Value& config);

    // kubectl apply / delete of a ComposabilityRequest
    Error Apply(const gojson::Value& request);
    Error Delete(const std::string& name);
    // test hook: plant a request/resource in a given state without running anything
    Error Plant(const gojson::Value& object);
    // runs both controllers until the cluster is quiescent (or max_reconciles)
    void Run(long long max_reconciles);
    void traceEvent(char what, int dev) {
        if (!stats.tracing || stats.trace.size() >= 4000) return;
        const long long t = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
        if (stats.trace.empty()) trace_t0_ = t;
        stats.trace.push_back({t - trace_t0_, what, dev});
    }
    long long trace_t0_ = 0;
    std::chrono::steady_clock::time_point last_poll_{};
    // exactly one Reconcile of the request controller on `name` (the reference's tests drive it this way)
    Error ReconcileRequestOnce(const std::string& name);
    Error ReconcileResourceOnce(const std::string& name);
    // one tick of the UpstreamSyncer (upstreamsyncer_controller.go:77-136) at time now_s
    Error SyncUpstream(const gojson::Value& devices, long long now_s);

    std::string DumpJSON() const;
    std::string StatsJSON() const;
    Stats stats;

private:
    friend class RequestReconciler;
    friend class ResourceReconciler;
    friend class SimProvider;
    friend class SimNodeOps;

    // ---- store with k8s semantics -----------------------------------------
    const Node* getNode(const std::string& name) const;
    void updateRequest(const ComposabilityRequest& r);          // Update + Status().Update
    void deleteRequest(const std::string& name);
    void createResource(const StoredResource& r);
    void updateResource(const StoredResource& r);
    void deleteResource(const std::string& name);
    void enqueueRequest(const std::string& key);
    void enqueueResource(const std::string& key);
    void enqueueResourceFront(const std::string& key);
    std::string GenerateComposableResourceName(const std::string& typeName);
    Error CheckNodeCapacitySufficient(const std::string& nodeName, const NodeSpec& spec, bool* ok) const;

    void pollProbes(bool block);
    void releaseDevice(int dev);
    Error reconcileRequest(const std::string& key, long long* requeue_after_s);
    Error reconcileResource(const std::string& key, long long* requeue_after_s);

    cro_ctx* ctx_;
    std::string deviceResourceType_ = "DEVICE_PLUGIN";
    // API-server fault injection (the reference's MockUpdate / MockStatusUpdate hooks): while armed,
    // Update() resp. Status().Update() of either kind fails with this text and writes nothing.
    std::string fault_update_, fault_status_update_;
    bool probe_ = false;
    std::vector<Node> nodes_;                                   // sorted by name (API list order)
    std::map<std::string, ComposabilityRequest> requests_;
    std::map<std::string, StoredResource> resources_;
    std::vector<std::string> uuids_;
    std::set<std::string> attached_;                            // resource names the fake fabric has attached
    std::map<std::string, long long> missing_devices_;          // UpstreamSyncer.missingDevices: id -> first seen (s)
    std::map<std::string, long long> cdi_serial_;               // per request: next res-<req>-<k>
    std::deque<std::string> req_queue_, res_queue_;
    std::set<std::string> req_queued_, res_queued_;
    std::set<std::string> req_timers_, res_timers_;
    std::map<int, std::deque<std::string>> probe_owner_;        // device -> the attaches whose probes are in flight on it, oldest
                                                                // first (up to two: the second one's kernels are queued behind
                                                                // the first's, so the GPU never waits for the host)
    std::set<int> probe_notified_;                              // devices whose owner has been told "done"
    std::map<int, std::deque<std::string>> dev_waiters_;        // attaches queued behind a device's owner
    std::set<std::string> probe_waiting_;                       // names present in some dev_waiters_ queue
    long long changes_ = 0, seq_ = 0;
    std::mt19937_64 rng_;
};

}  // namespace sim
}  // namespace cro
