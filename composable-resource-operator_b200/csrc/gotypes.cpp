// gotypes.cpp — see gotypes.hpp.  Field lists follow the struct declarations (json tags), types follow
// reflect.Type.String() ("api." is the package name of both flavours).
#include "gotypes.hpp"

#include <deque>

namespace cro {
namespace gotypes {
namespace {

using gojson::GoType;
typedef std::vector<std::pair<std::string, const GoType*>> Fields;

std::deque<GoType>& pool() {
    static std::deque<GoType> p;       // deque: addresses stay put
    return p;
}
const GoType* prim(GoType::Kind k, const char* name) {
    pool().push_back(GoType());
    GoType& t = pool().back();
    t.kind = k;
    t.name = name;
    return &t;
}
const GoType* strct(const char* name, const Fields& fields) {
    pool().push_back(GoType());
    GoType& t = pool().back();
    t.kind = GoType::Struct;
    t.name = std::string("api.") + name;
    t.structName = name;
    t.fields = fields;
    return &t;
}
const GoType* slice(const GoType* elem) {
    pool().push_back(GoType());
    GoType& t = pool().back();
    t.kind = GoType::Slice;
    t.name = "[]" + elem->name;
    t.elem = elem;
    return &t;
}

struct All {
    const GoType *S, *I, *B;
    const GoType *fmScaleUp, *fmGetMachine, *cmMachineData;
    All() {
        S = prim(GoType::String, "string");
        I = prim(GoType::Int, "int");
        B = prim(GoType::Bool, "bool");
        // fm/api/common.go:19-29
        const GoType* item = strct("ConditionItem", {{"column", S}, {"operator", S}, {"value", S}});
        const GoType* cond = strct("Condition", {{"condition", slice(item)}});
        // fm/api/scale_up.go:43-69
        const GoType* upRes = strct("ScaleUpResponseResourceItem", {{"res_uuid", S}, {"res_name", S}, {"res_type", S}, {"res_status", I},
                                                                    {"res_op_status", S}, {"res_serial_num", S}, {"res_spec", cond}});
        const GoType* upMach = strct("ScaleUpResponseMachineItem", {{"fabric_uuid", S}, {"fabric_id", I}, {"mach_uuid", S}, {"mach_id", I},
                                                                    {"mach_name", S}, {"tenant_uuid", S}, {"resources", slice(upRes)}});
        fmScaleUp = strct("ScaleUpResponse", {{"data", strct("ScaleUpResponseData", {{"machines", slice(upMach)}})}});
        // fm/api/get.go:19-49
        const GoType* res = strct("GetMachineResource", {{"res_uuid", S}, {"res_name", S}, {"res_type", S}, {"res_status", I},
                                                         {"res_op_status", S}, {"res_serial_num", S}, {"res_spec", cond}});
        const GoType* mach = strct("GetMachineItem", {{"fabric_uuid", S}, {"fabric_id", I}, {"mach_uuid", S}, {"mach_id", I}, {"mach_name", S},
                                                      {"tenant_uuid", S}, {"mach_status", I}, {"mach_status_detail", S}, {"resources", slice(res)}});
        fmGetMachine = strct("GetMachineResponse", {{"data", strct("GetMachineData", {{"machines", slice(mach)}})}});
        // cm/api/machine.go:19-93
        const GoType* cmCond = strct("Condition", {{"column", S}, {"operator", S}, {"value", S}});
        const GoType* selector = strct("Selector", {{"version", S}, {"expression", strct("Expression", {{"conditions", slice(cmCond)}})}});
        const GoType* devSpec = strct("DeviceResourceSpec", {{"resspec_uuid", S}, {"productname", S}, {"model", S}, {"vendor", S}, {"removable", B}});
        const GoType* detail = strct("DeviceDetail", {{"fabric_uuid", S}, {"fabric_id", I}, {"res_uuid", S}, {"fabr_gid", S}, {"res_type", S},
                                                      {"res_name", S}, {"res_status", S}, {"res_op_status", S}, {"resspecs", slice(devSpec)},
                                                      {"tenant_uuid", S}, {"mach_uuid", S}});
        const GoType* device = strct("Device", {{"device_id", S}, {"status", S}, {"status_reason", S}, {"detail", detail}});
        const GoType* spec = strct("ResourceSpec", {{"spec_uuid", S}, {"type", S}, {"selector", selector}, {"min_resspec_count", I},
                                                    {"max_resspec_count", I}, {"device_count", I}, {"devices", slice(device)}});
        const GoType* machine = strct("Machine", {{"uuid", S}, {"name", S}, {"status", S}, {"status_reason", S}, {"resspecs", slice(spec)}});
        const GoType* cluster = strct("Cluster", {{"cluster_uuid", S}, {"machine", machine}});
        cmMachineData = strct("MachineData", {{"data", strct("Data", {{"tenant_uuid", S}, {"cluster", cluster}})}});
    }
};
const All& all() {
    static const All a;
    return a;
}

}  // namespace

const gojson::GoType& FMScaleUpResponse() { return *all().fmScaleUp; }
const gojson::GoType& FMGetMachineResponse() { return *all().fmGetMachine; }
const gojson::GoType& CMMachineData() { return *all().cmMachineData; }

}  // namespace gotypes
}  // namespace cro
