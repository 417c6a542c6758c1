// fabric.cpp — see fabric.hpp.
#include "fabric.hpp"
#include "gotypes.hpp"

#include "gojson.hpp"

namespace cro {
namespace fabric {

namespace {
using gojson::Value;

const Value* arr(const Value* v, const char* k) {
    const Value* a = v ? v->get(k) : nullptr;
    return (a && a->kind == Value::Array) ? a : nullptr;
}

// res_op_status[:1] decision shared by the four call sites; `where` is "FM" or "CM".
Error opStatusDecision(const std::string& op, const std::string& deviceID, const char* where) {
    if (op.empty()) return Error::New("runtime error: slice bounds out of range [:1] with length 0");
    if (op[0] == '0') return Error::Nil();
    if (op[0] == '1') return Error::New("the target gpu '" + deviceID + "' is showing a Warning status in " + where);
    if (op[0] == '2') return Error::New("the target gpu '" + deviceID + "' is showing a Critical status in " + where);
    return Error::New("the target gpu '" + deviceID + "' has unknown status '" + op + "' in " + where);
}

Error notFound(const std::string& deviceID) {
    return Error::New("the target device '" + deviceID + "' cannot be found in CDI system");
}
}  // namespace

Error FMCheckResource(const std::string& body, const std::string& specType, const std::string& specModel,
                      const std::string& deviceID) {
    std::string perr;
    gojson::ValuePtr root = gojson::DecodeAs(gojson::parse(body, &perr), body, gotypes::FMGetMachineResponse(), &perr);
    if (!root)
        return Error::New("failed to unmarshal FM get machine response body into machineData: " + perr);
    const Value* machines = arr(root->get("data"), "machines");
    if (!machines || machines->arr.empty())   // fm/client.go:331 indexes Machines[0] unguarded
        return Error::New("runtime error: index out of range [0] with length 0");
    const Value* resources = arr(machines->arr[0].get(), "resources");
    if (resources)
        for (const auto& r : resources->arr) {
            if (r->get_string("res_type") != specType) continue;
            const Value* conds = arr(r->get("res_spec"), "condition");
            if (!conds) continue;
            for (const auto& c : conds->arr) {
                if (c->get_string("column") != "model" || c->get_string("operator") != "eq" ||
                    c->get_string("value") != specModel)
                    continue;
                if (r->get_string("res_serial_num") == deviceID)
                    return opStatusDecision(r->get_string("res_op_status"), deviceID, "FM");
            }
        }
    return notFound(deviceID);
}

Error FMGetResources(const std::string& body, const std::string& nodeName, const std::string& machineID,
                     std::vector<DeviceInfo>* out) {
    std::string perr;
    gojson::ValuePtr root = gojson::DecodeAs(gojson::parse(body, &perr), body, gotypes::FMGetMachineResponse(), &perr);
    if (!root)
        return Error::New("failed to unmarshal FM get machine response body into machineData: " + perr);
    const Value* machines = arr(root->get("data"), "machines");
    if (!machines || machines->arr.empty()) return Error::Nil();   // fm/client.go:385-387
    const Value* resources = arr(machines->arr[0].get(), "resources");
    if (!resources) return Error::Nil();
    for (const auto& r : resources->arr) {
        if (r->get_string("res_type") != "gpu") continue;
        std::string model;
        if (const Value* conds = arr(r->get("res_spec"), "condition"))
            for (const auto& c : conds->arr)
                if (c->get_string("column") == "model" && c->get_string("operator") == "eq") {
                    model = c->get_string("value");
                    break;
                }
        out->push_back({nodeName, machineID, r->get_string("res_type"), model, r->get_string("res_serial_num"),
                        r->get_string("res_uuid")});
    }
    return Error::Nil();
}

Error CMCheckResource(const std::string& body, const std::string& specType, const std::string& specModel,
                      const std::string& deviceID) {
    std::string perr;
    gojson::ValuePtr root = gojson::DecodeAs(gojson::parse(body, &perr), body, gotypes::CMMachineData(), &perr);
    if (!root)
        return Error::New("failed to unmarshal CM get machine response body into machineData: " + perr);
    const Value* data = root->get("data");
    const Value* cluster = data ? data->get("cluster") : nullptr;
    const Value* specs = arr(cluster ? cluster->get("machine") : nullptr, "resspecs");
    if (specs)
        for (const auto& s : specs->arr) {
            if (s->get_string("type") != specType) continue;
            const Value* sel = s->get("selector");
            const Value* conds = arr(sel ? sel->get("expression") : nullptr, "conditions");
            if (!conds) continue;
            for (const auto& c : conds->arr) {
                if (c->get_string("column") != "model" || c->get_string("operator") != "eq" ||
                    c->get_string("value") != specModel)
                    continue;
                if (const Value* devs = arr(s.get(), "devices"))
                    for (const auto& d : devs->arr)
                        if (d->get_string("device_id") == deviceID) {
                            const Value* detail = d->get("detail");
                            return opStatusDecision(detail ? detail->get_string("res_op_status") : std::string(), deviceID, "CM");
                        }
            }
        }
    return notFound(deviceID);
}

Error CMGetResources(const std::string& body, const std::string& nodeName, const std::string& machineID,
                     std::vector<DeviceInfo>* out) {
    std::string perr;
    gojson::ValuePtr root = gojson::DecodeAs(gojson::parse(body, &perr), body, gotypes::CMMachineData(), &perr);
    if (!root)
        return Error::New("failed to unmarshal CM get machine response body into machineData: " + perr);
    const Value* data = root->get("data");
    const Value* cluster = data ? data->get("cluster") : nullptr;
    const Value* specs = arr(cluster ? cluster->get("machine") : nullptr, "resspecs");
    if (!specs) return Error::Nil();
    for (const auto& s : specs->arr) {
        if (s->get_string("type") != "gpu") continue;
        if (const Value* devs = arr(s.get(), "devices"))
            for (const auto& d : devs->arr) {
                const Value* detail = d->get("detail");
                // cm/client.go:335-341: Model is left empty by the CM flavour
                out->push_back({nodeName, machineID, s->get_string("type"), std::string(), d->get_string("device_id"),
                                detail ? detail->get_string("res_uuid") : std::string()});
            }
    }
    return Error::Nil();
}

std::string DeviceInfosToJson(const std::vector<DeviceInfo>& v) {
    gojson::Writer w;
    w.begin_array();
    for (const DeviceInfo& d : v) {
        w.begin_object();
        w.field("node_name", d.NodeName).field("machine_uuid", d.MachineUUID).field("device_type", d.DeviceType);
        w.field("model", d.Model).field("device_id", d.DeviceID).field("cdi_device_id", d.CDIDeviceID);
        w.end_object();
    }
    w.end_array();
    return w.take();
}

}  // namespace fabric
}  // namespace cro
