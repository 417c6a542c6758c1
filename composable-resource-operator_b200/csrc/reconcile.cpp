// reconcile.cpp — see reconcile.hpp.
#include "reconcile.hpp"
#include "gotypes.hpp"

namespace cro {
namespace controller {

const std::string ErrWaitingDeviceAttaching = "device is attaching to the cluster";
const std::string ErrWaitingDeviceDetaching = "device is detaching from the cluster";

std::string ComposableResourceStatus::MarshalJSON() const {
    // api/v1alpha1/composableresource_types.go:36-41 — `state` is always
    // present, the other three are omitempty.
    gojson::Writer w;
    w.begin_object();
    w.field("state", State);
    w.field_omitempty("error", Error);
    w.field_omitempty("device_id", DeviceID);
    w.field_omitempty("cdi_device_id", CDIDeviceID);
    w.end_object();
    return w.take();
}

Error FMScaleUpResponseToIDs(const std::string& body, const std::string& instanceName,
                             const std::string& specType, const std::string& specModel,
                             std::string* deviceID, std::string* CDIDeviceID) {
    deviceID->clear();
    CDIDeviceID->clear();
    std::string perr;
    gojson::ValuePtr root = gojson::DecodeAs(gojson::parse(body, &perr), body, gotypes::FMScaleUpResponse(), &perr);
    if (!root)
        return Error::New(
            "failed to unmarshal FM scaleup response body into scaleUpResponse. Original error: " + perr);
    const gojson::Value* data = root->get("data");
    const gojson::Value* machines = data ? data->get("machines") : nullptr;
    if (machines && machines->kind == gojson::Value::Array && !machines->arr.empty()) {
        const gojson::Value* m0 = machines->arr[0].get();
        const gojson::Value* resources = m0->get("resources");
        if (resources && resources->kind == gojson::Value::Array && !resources->arr.empty()) {
            const gojson::Value* r0 = resources->arr[0].get();
            if (r0->get_string("res_type") == specType) {
                const gojson::Value* spec = r0->get("res_spec");
                const gojson::Value* cond = spec ? spec->get("condition") : nullptr;
                if (cond && cond->kind == gojson::Value::Array) {
                    for (const auto& it : cond->arr) {
                        if (it->get_string("column") == "model" && it->get_string("operator") == "eq" &&
                            it->get_string("value") == specModel) {
                            const std::string op = r0->get_string("res_op_status");
                            if (op.empty())  // Go: OptionStatus[:1] on "" panics
                                return Error::New("runtime error: slice bounds out of range [:1] with length 0");
                            const char c = op[0];
                            if (c == '0' || c == '1') {
                                *deviceID = r0->get_string("res_serial_num");
                                *CDIDeviceID = r0->get_string("res_uuid");
                                return Error::Nil();
                            }
                            if (c == '2')
                                return Error::New("the FM attached device called by " + instanceName +
                                                  " is in Critical state in FM");
                            return Error::New("the FM attached device called by " + instanceName +
                                              " is in unknown state '" + op + "' in FM");
                        }
                    }
                }
            }
        }
    }
    return Error::New("can not find the added gpu when using FM to add gpu");
}

CMAddingResult CMCheckAddingResources(const std::string& machineBody,
                                      const std::vector<std::string>& existingDeviceIDs,
                                      const std::string& specType, const std::string& specModel) {
    CMAddingResult out;
    std::string perr;
    gojson::ValuePtr root = gojson::DecodeAs(gojson::parse(machineBody, &perr), machineBody, gotypes::CMMachineData(), &perr);
    if (!root) {
        out.err = Error::New("failed to unmarshal CM get machine response body into machineData: " + perr);
        return out;
    }
    const gojson::Value* data = root->get("data");
    const gojson::Value* cluster = data ? data->get("cluster") : nullptr;
    const gojson::Value* machine = cluster ? cluster->get("machine") : nullptr;
    const gojson::Value* specs = machine ? machine->get("resspecs") : nullptr;
    if (!specs || specs->kind != gojson::Value::Array) return out;
    for (const auto& spec : specs->arr) {
        // isSpecMatch (cm/client.go:485-499)
        if (spec->get_string("type") != specType) continue;
        bool match = false;
        const gojson::Value* sel = spec->get("selector");
        const gojson::Value* expr = sel ? sel->get("expression") : nullptr;
        const gojson::Value* conds = expr ? expr->get("conditions") : nullptr;
        if (conds && conds->kind == gojson::Value::Array)
            for (const auto& c : conds->arr)
                if (c->get_string("column") == "model" && c->get_string("operator") == "eq" &&
                    c->get_string("value") == specModel)
                    match = true;
        if (!match) continue;
        // findAvailableDevice (cm/client.go:501-509): first device no CR owns
        const gojson::Value* devices = spec->get("devices");
        if (devices && devices->kind == gojson::Value::Array) {
            for (const auto& dev : devices->arr) {
                const std::string id = dev->get_string("device_id");
                bool owned = false;
                for (const std::string& e : existingDeviceIDs)
                    if (e == id) owned = true;
                if (owned) continue;
                const std::string status = dev->get_string("status");
                const gojson::Value* detail = dev->get("detail");
                const std::string res = detail ? detail->get_string("res_uuid") : std::string();
                if (status == "ADD_COMPLETE") {
                    out.deviceID = id;
                    out.CDIDeviceID = res;
                    return out;
                }
                if (status == "ADD_FAILED") {
                    out.deviceID = id;
                    out.CDIDeviceID = res;
                    out.err = Error::New("an error occurred with the resource in CM: '" +
                                         dev->get_string("status_reason") + "'");
                    return out;
                }
                break;  // an unowned device in any other state: fall through to the resize request
            }
        }
        out.specUUID = spec->get_string("spec_uuid");
        out.deviceCount = spec->get_int("device_count");
        break;
    }
    return out;
}

// A callee panicked: unwind at once, no status write on the way (see Error::panicked).
#define CRO_UNWIND(e) do { if ((e).panicked() || (e).recovered()) return Error::Recovered(e); } while (0)
// `if err := r.Status().Update(ctx, resource); err != nil { return r.requeueOnErr(resource, err, ...) }`
#define CRO_UPDATE_OR_REQUEUE(resource) do { Error ue__ = statusUpdate(*(resource)); if (!ue__.ok()) return requeueOnErr((resource), ue__); } while (0)

Error ComposableResourceReconciler::requeueOnErr(ComposableResource* resource, const Error& err) {
    CRO_UNWIND(err);                       // a panic never reaches requeueOnErr in the reference
    if (resource) {
        resource->Status.Error = err.msg;
        (void)statusUpdate(*resource);     // :428-430: a failure of THIS write is only logged
    }
    return err;
}

Error ComposableResourceReconciler::handleNoneState(ComposableResource* resource, Result* result) {
    *result = Result();
    auto it = resource->Labels.find("cohdi.io/ready-to-detach-device-id");
    if (it != resource->Labels.end() && !it->second.empty()) {
        resource->Status.DeviceID = it->second;
        auto jt = resource->Labels.find("cohdi.io/ready-to-detach-cdi-device-id");
        if (jt != resource->Labels.end() && !jt->second.empty()) resource->Status.CDIDeviceID = jt->second;
    }
    resource->Status.State = "Attaching";
    resource->Status.Error = "";
    return statusUpdate(*resource);        // :197 `return ctrl.Result{}, r.Status().Update(ctx, resource)`
}

Error ComposableResourceReconciler::handleAttachingState(ComposableResource* resource,
                                                         const std::string& deviceResourceType,
                                                         Result* result) {
    *result = Result();
    if (resource->DeletionTimestampSet) {
        if (resource->Status.DeviceID.empty()) {
            resource->Status.State = "Deleting";
            return statusUpdate(*resource);            // :206
        }
        if (!resource->Status.Error.empty()) {
            resource->Status.State = "Detaching";
            return statusUpdate(*resource);            // :210
        }
        // DeviceID set, no error: keeps attaching
    }

    if (resource->Status.DeviceID.empty()) {
        std::string deviceID, cdiDeviceID;
        Error err = provider_->AddResource(*resource, &deviceID, &cdiDeviceID);
        CRO_UNWIND(err);                               // e.g. res_op_status[:1] on "" (fti/fm/client.go:195)
        if (!err.ok()) {
            if (err.msg == ErrWaitingDeviceAttaching) {  // errors.Is on the sentinel
                result->RequeueAfterSeconds = 30;
                return Error::Nil();
            }
            return requeueOnErr(resource, err);
        }
        resource->Status.Error = "";
        resource->Status.DeviceID = deviceID;
        resource->Status.CDIDeviceID = cdiDeviceID;
        CRO_UPDATE_OR_REQUEUE(resource);               // :233-235: nothing below runs when the IDs could not be stored
    }

    if (deviceResourceType == "DEVICE_PLUGIN") {
        Error le = node_->CheckNoGPULoads(resource->Spec.TargetNode);  // an error is logged only; a panic is not an error
        CRO_UNWIND(le);
        for (const char* ds : {"nvidia-device-plugin-daemonset", "nvidia-dcgm"}) {
            Error err = node_->RestartDaemonset("nvidia-gpu-operator", ds);
            CRO_UNWIND(err);
            if (!err.ok()) {
                resource->Status.Error = err.msg;
                CRO_UPDATE_OR_REQUEUE(resource);       // :247-250, :255-258
            }
        }
    } else if (deviceResourceType == "DRA") {
        Error err = node_->RunNvidiaSmi(resource->Spec.TargetNode);
        CRO_UNWIND(err);                               // e.g. parts[i] on a short CSV row (gpus.go:912-914)
        if (!err.ok()) {
            resource->Status.Error = err.msg;
            CRO_UPDATE_OR_REQUEUE(resource);           // :262-265
        }
        err = node_->RestartDaemonset("nvidia-dra-driver-gpu", "nvidia-dra-driver-gpu-kubelet-plugin");
        CRO_UNWIND(err);
        if (!err.ok()) {
            resource->Status.Error = err.msg;
            CRO_UPDATE_OR_REQUEUE(resource);           // :269-272
        }
    }

    bool visible = false;
    Error err = node_->CheckGPUVisible(deviceResourceType, *resource, &visible);
    CRO_UNWIND(err);
    if (!err.ok()) return requeueOnErr(resource, err);
    if (visible) {
        resource->Status.State = "Online";
        resource->Status.Error = "";
        return statusUpdate(*resource);                // :282
    }
    result->RequeueAfterSeconds = 30;
    return Error::Nil();
}

Error ComposableResourceReconciler::handleOnlineState(ComposableResource* resource, Result* result) {
    *result = Result();
    if (resource->DeletionTimestampSet) {   // :292-295
        resource->Status.State = "Detaching";
        return statusUpdate(*resource);
    }
    auto lb = resource->Labels.find("cohdi.io/ready-to-detach-device-id");
    if (lb != resource->Labels.end() && !lb->second.empty()) {   // :297-302
        resource->DeleteRequested = true;
        return Error::Nil();
    }
    Error err = provider_->CheckResource(*resource);   // :305-315: recorded, never returned
    CRO_UNWIND(err);                                    // ... unless it panicked (fti/fm/client.go:346, cm/client.go:291)
    resource->Status.Error = err.ok() ? std::string() : err.msg;
    CRO_UPDATE_OR_REQUEUE(resource);
    result->RequeueAfterSeconds = 30;
    return Error::Nil();
}

Error ComposableResourceReconciler::handleDetachingState(ComposableResource* resource,
                                                         const std::string& deviceResourceType,
                                                         Result* result) {
    *result = Result();
    if (!resource->Status.DeviceID.empty()) {
        if (!resource->Spec.ForceDetach) {   // :327-341
            Error err = deviceResourceType == "DEVICE_PLUGIN"
                            ? node_->CheckNoGPULoadsFor(resource->Spec.TargetNode, nullptr)
                            : node_->CheckNoGPULoadsFor(resource->Spec.TargetNode, &resource->Status.DeviceID);
            CRO_UNWIND(err);
            if (!err.ok()) return requeueOnErr(resource, err);
        }
        if (deviceResourceType == "DRA") {   // :344-348
            Error err = node_->CreateDeviceTaint(*resource);
            if (!err.ok()) return requeueOnErr(resource, err);
        }
        Error err = node_->DrainGPU(resource->Spec.TargetNode, resource->Status.DeviceID, deviceResourceType);   // :351
        if (!err.ok()) return requeueOnErr(resource, err);
        err = provider_->RemoveResource(*resource);   // :355-364
        if (!err.ok()) {
            if (err.msg == ErrWaitingDeviceDetaching) {
                result->RequeueAfterSeconds = 30;
                return Error::Nil();
            }
            return requeueOnErr(resource, err);
        }
        if (deviceResourceType == "DEVICE_PLUGIN") {   // :369-380: restart failures are fatal here
            for (const char* ds : {"nvidia-device-plugin-daemonset", "nvidia-dcgm"}) {
                err = node_->RestartDaemonset("nvidia-gpu-operator", ds);
                if (!err.ok()) return requeueOnErr(resource, err);
            }
        } else {
            err = node_->RestartDaemonset("nvidia-dra-driver-gpu", "nvidia-dra-driver-gpu-kubelet-plugin");
            if (!err.ok()) return requeueOnErr(resource, err);
        }
        bool visible = false;   // :383-390
        err = node_->CheckGPUVisible(deviceResourceType, *resource, &visible);
        if (!err.ok()) return requeueOnErr(resource, err);
        if (visible) {
            result->RequeueAfterSeconds = 3;
            return Error::Nil();
        }
        if (deviceResourceType == "DRA") {   // :393-397
            err = node_->DeleteDeviceTaint(*resource);
            if (!err.ok()) return requeueOnErr(resource, err);
        }
        resource->Status.Error = "";
        resource->Status.DeviceID = "";
        resource->Status.CDIDeviceID = "";
        CRO_UPDATE_OR_REQUEUE(resource);   // :401-403
    }
    resource->Status.State = "Deleting";   // :405-406
    return statusUpdate(*resource);
}

}  // namespace controller
}  // namespace cro
