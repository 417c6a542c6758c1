// provider.cpp — see provider.hpp.
#include "provider.hpp"
#include "gotypes.hpp"

#include "identity.hpp"

namespace cro {
namespace fabric {

using controller::ComposableResource;
using controller::ErrWaitingDeviceAttaching;
using controller::ErrWaitingDeviceDetaching;
using gojson::Value;

// ---------------------------------------------------------------------------
// request bodies
// ---------------------------------------------------------------------------
std::string FMScaleUpBody(const std::string& tenant, const std::string& machine, const std::string& type,
                          const std::string& model) {
    gojson::Writer w;
    w.begin_object().key("tenants").begin_object();
    w.field("tenant_uuid", tenant);
    w.key("machines").begin_array().begin_object();
    w.field("mach_uuid", machine);
    w.key("resources").begin_array().begin_object();
    w.key("res_specs").begin_array().begin_object();
    w.field("res_type", type);
    w.key("res_spec").begin_object().key("condition").begin_array().begin_object();
    w.field("column", std::string("model")).field("operator", std::string("eq")).field("value", model);
    w.end_object().end_array().end_object();
    w.field("res_num", 1);
    w.end_object().end_array();      // res_specs
    w.end_object().end_array();      // resources
    w.end_object().end_array();      // machines
    w.end_object().end_object();
    return w.str();
}

std::string FMScaleDownBody(const std::string& tenant, const std::string& machine, const std::string& type,
                            const std::string& resUUID) {
    gojson::Writer w;
    w.begin_object().key("tenants").begin_object();
    w.field("tenant_uuid", tenant);
    w.key("machines").begin_array().begin_object();
    w.field("mach_uuid", machine);
    w.key("resources").begin_array().begin_object();
    w.key("res_specs").begin_array().begin_object();
    w.field("res_type", type);
    w.field("res_uuid", resUUID);
    w.field("res_num", 1);
    w.end_object().end_array();
    w.end_object().end_array();
    w.end_object().end_array();
    w.end_object().end_object();
    return w.str();
}

std::string CMScaleUpBody(const std::string& specUUID, long long deviceCount) {
    gojson::Writer w;
    w.begin_object().key("increase_resource_count").begin_object();
    w.field("spec_uuid", specUUID).field("device_count", deviceCount);
    w.end_object().end_object();
    return w.str();
}

std::string CMScaleDownBody(const std::string& specUUID, long long deviceCount, const std::string& device) {
    gojson::Writer w;
    w.begin_object().key("remove_resources").begin_object();
    w.field("spec_uuid", specUUID).field("device_count", deviceCount);
    w.key("devices").begin_array().value(device).end_array();
    w.end_object().end_object();
    return w.str();
}

// ---------------------------------------------------------------------------
// ErrorBody decoding
// ---------------------------------------------------------------------------
namespace {

const char* goKind(const Value& v) {   // the word json.UnmarshalTypeError.Value carries
    switch (v.kind) {
        case Value::Object: return "object";
        case Value::Array: return "array";
        case Value::String: return "string";
        case Value::Number: return "number";
        case Value::Bool: return "bool";
        default: return "null";
    }
}

// Unmarshal(body, &api.ErrorBody{}) for both flavours.  Go keeps decoding after a
// type mismatch and reports the FIRST one; fields that do decode keep their values.
// (Type-mismatch texts follow go1.24's UnmarshalTypeError.Error(); the reference
// pins only the syntax-error form — parity of the mismatch form is unpinned.)
struct ErrorBody {
    long long status = 0;
    std::string code;
    std::string message;            // CM: detail.message as a string
    const Value* rawMessage = nullptr;   // FM: detail.message as json.RawMessage
    const Value* detail = nullptr;
    gojson::ValuePtr root;
};

std::string decodeErrorBody(const std::string& body, bool fm, ErrorBody* out) {
    std::string perr;
    out->root = gojson::parse(body, &perr);
    if (!out->root) return perr;
    const Value& root = *out->root;
    if (root.kind == Value::Null) return "";
    if (root.kind != Value::Object)
        return std::string("json: cannot unmarshal ") + goKind(root) + " into Go value of type api.ErrorBody";
    // members in input order, as the decoder walks them: the FIRST mismatch is the one Unmarshal returns, and a
    // later well-typed duplicate still overwrites the field
    std::string first;
    auto mismatch = [&](const Value& v, const char* where, const char* type, bool intTarget) {
        if (first.empty())
            first = "json: cannot unmarshal " + gojson::MismatchKind(v, body, intTarget) + " into Go struct field " + where + " of type " + type;
    };
    static const char* const kBody[] = {"status", "detail"};
    static const char* const kDetail[] = {"code", "message", "data"};
    for (const auto& kv : root.obj) {
        const int f = gojson::MatchField(kv.first, kBody, 2);
        const Value& v = *kv.second;
        if (f < 0 || v.kind == Value::Null) continue;
        if (f == 0) {
            if (v.kind == Value::Number && v.is_int) out->status = v.inum;
            else mismatch(v, "ErrorBody.status", "int", true);
            continue;
        }
        if (v.kind != Value::Object) { mismatch(v, "ErrorBody.detail", "api.ErrorDetail", false); continue; }
        out->detail = &v;
        for (const auto& dk : v.obj) {
            const int g = gojson::MatchField(dk.first, kDetail, fm ? 3 : 2);
            const Value& m = *dk.second;
            if (g < 0) continue;
            if (g == 1 && fm) { out->rawMessage = &m; continue; }         // json.RawMessage takes any value, null included
            if (m.kind == Value::Null) continue;
            if (g == 2) {
                if (m.kind != Value::Object) mismatch(m, "ErrorDetail.detail.data", "map[string]interface {}", false);
            } else if (m.kind == Value::String) {
                (g == 0 ? out->code : out->message) = m.str;
            } else {
                mismatch(m, g == 0 ? "ErrorDetail.detail.code" : "ErrorDetail.detail.message", "string", false);
            }
        }
    }
    return first;
}

}  // namespace

std::string formatFMErrorDetail(const std::string& code, const Value* m, const std::string& text) {
    std::string message;
    if (m) {
        if (m->kind == Value::String) message = m->str;                 // it unmarshals into a Go string
        else if (m->kind == Value::Null) message = "";                  // Unmarshal("null", &s) leaves s == ""
        else message = identity::TrimSpace(text.substr(m->raw_begin, m->raw_end - m->raw_begin));
    }
    return "code: '" + code + "', error message: '" + message + "'";
}

Error FMErrorFromReply(const std::string& what, const std::string& body) {
    ErrorBody eb;
    const std::string uerr = decodeErrorBody(body, true, &eb);
    if (!uerr.empty()) {
        // the scaledown flavour drops the "FM" word (fm/client.go:303)
        const std::string subject = what == "scaledown" ? "scaledown" : "FM " + what;
        return Error::New("failed to unmarshal " + subject + " error response body into errBody. Original error: " + uerr);
    }
    return Error::New("failed to process FM " + what + " request. FM returned " + formatFMErrorDetail(eb.code, eb.rawMessage, body));
}

Error CMErrorFromReply(const std::string& what, const std::string& body) {
    ErrorBody eb;
    const std::string uerr = decodeErrorBody(body, false, &eb);
    if (!uerr.empty())
        return Error::New("failed to unmarshal CM " + what + " error response body into errBody. Original error: " + uerr);
    if (what == "scaledown")   // no quotes in this one (cm/client.go:255; Appendix A-9)
        return Error::New("failed to process CM scaledown request. http returned status: " + std::to_string(eb.status) +
                          ", cm return code: " + eb.code + ", error message: " + eb.message);
    return Error::New("failed to process CM " + what + " request. http returned status: '" + std::to_string(eb.status) +
                      "', cm return code: '" + eb.code + "', error message: '" + eb.message + "'");
}

// ---------------------------------------------------------------------------
// token reply
// ---------------------------------------------------------------------------
bool DecodeBase64RawURL(const std::string& in, std::string* out, std::string* err) {
    // encoding/base64 decodeQuantum for an unpadded URL alphabet: CR / LF are skipped, any other byte outside the
    // alphabet is corrupt at its own offset, and a dangling single character is corrupt at ITS offset (si - j)
    auto val = [](unsigned char c) -> int {
        if (c >= 'A' && c <= 'Z') return c - 'A';
        if (c >= 'a' && c <= 'z') return c - 'a' + 26;
        if (c >= '0' && c <= '9') return c - '0' + 52;
        if (c == '-') return 62;
        if (c == '_') return 63;
        return -1;
    };
    out->clear();
    unsigned acc[4];
    int j = 0;
    size_t si = 0;
    for (; si < in.size(); ++si) {
        const unsigned char c = (unsigned char)in[si];
        if (c == '\n' || c == '\r') continue;
        const int v = val(c);
        if (v < 0) {
            *err = "illegal base64 data at input byte " + std::to_string(si);
            return false;
        }
        acc[j++] = (unsigned)v;
        if (j == 4) {
            out->push_back((char)((acc[0] << 2) | (acc[1] >> 4)));
            out->push_back((char)(((acc[1] & 15) << 4) | (acc[2] >> 2)));
            out->push_back((char)(((acc[2] & 3) << 6) | acc[3]));
            j = 0;
        }
    }
    if (j == 1) {
        *err = "illegal base64 data at input byte " + std::to_string(si - 1);
        return false;
    }
    if (j >= 2) out->push_back((char)((acc[0] << 2) | (acc[1] >> 4)));
    if (j == 3) out->push_back((char)(((acc[1] & 15) << 4) | (acc[2] >> 2)));
    return true;
}

Error TokenFromReply(const TokenReply& r, long long* expiryUnix) {
    if (!r.secret_error.empty()) return Error::New(r.secret_error);
    if (!r.transport_error.empty()) return Error::New(r.transport_error);
    if (r.status != 200) return Error::New("http returned code: " + std::to_string(r.status) + ", response body: " + r.body);
    std::string perr;
    gojson::ValuePtr root = gojson::parse(r.body, &perr);
    if (!gojson::rootOk(root, &perr, "fti.token"))
        return Error::New("failed to read id_manager response body into Token: " + perr);
    static const gojson::FlatField kToken[] = {{"access_token", 's'}, {"expires_in", 'i'}, {"refresh_expires_in", 'i'},
                                               {"refresh_token", 's'}, {"token_type", 's'}, {"id_token", 's'},
                                               {"not-before-policy", 'i'}, {"session_state", 's'}, {"scope", 's'}};
    static const gojson::FlatField kClaims[] = {{"exp", 'i'}};
    std::map<std::string, std::string> strs;
    std::map<std::string, long long> ints;
    perr = gojson::DecodeFlat(*root, r.body, "token", kToken, sizeof kToken / sizeof kToken[0], &strs, &ints);
    if (!perr.empty()) return Error::New("failed to read id_manager response body into Token: " + perr);
    const std::string access = strs["access_token"];
    const std::vector<std::string> parts = identity::Split(access, ".");
    if (parts.size() != 3) return Error::New("invalid access token: " + access);
    std::string payload, derr;
    if (!DecodeBase64RawURL(parts[1], &payload, &derr)) return Error::New("failed to decode id_manager payload: " + derr);
    gojson::ValuePtr claims = gojson::parse(payload, &perr);
    if (!gojson::rootOk(claims, &perr, "fti.accessToken")) return Error::New("failed to unmarshal id_manager json: " + perr);
    strs.clear();
    ints.clear();
    perr = gojson::DecodeFlat(*claims, payload, "accessToken", kClaims, 1, &strs, &ints);
    if (!perr.empty()) return Error::New("failed to unmarshal id_manager json: " + perr);
    *expiryUnix = ints["exp"];
    return Error::Nil();
}

Error ReplyTokenSource::GetToken() {
    if (have_ && expiry_ - 30 > now_) return Error::Nil();            // leeway 30 s (token.go:68,78)
    ++fetches;
    long long exp = 0;
    Error e = TokenFromReply(reply_, &exp);
    if (!e.ok()) return Error::New("unable to rotate token: " + e.msg);
    have_ = true;
    expiry_ = exp;
    return Error::Nil();
}

// ---------------------------------------------------------------------------
// CM checkRemovingResources
// ---------------------------------------------------------------------------
namespace {
const Value* arrOf(const Value* v, const char* k) {
    const Value* a = v ? v->get(k) : nullptr;
    return (a && a->kind == Value::Array) ? a : nullptr;
}
const Value* cmResspecs(const Value* root) {
    const Value* data = root ? root->get("data") : nullptr;
    const Value* cluster = data ? data->get("cluster") : nullptr;
    return arrOf(cluster ? cluster->get("machine") : nullptr, "resspecs");
}
bool cmSpecMatch(const Value& spec, const std::string& type, const std::string& model) {   // cm/client.go:485-499
    if (spec.get_string("type") != type) return false;
    const Value* sel = spec.get("selector");
    if (const Value* conds = arrOf(sel ? sel->get("expression") : nullptr, "conditions"))
        for (const auto& c : conds->arr)
            if (c->get_string("column") == "model" && c->get_string("operator") == "eq" && c->get_string("value") == model)
                return true;
    return false;
}
}  // namespace

CMRemovingResult CMCheckRemovingResources(const std::string& machineBody, const std::string& specType,
                                          const std::string& specModel, const std::string& deviceID) {
    CMRemovingResult r;
    std::string perr;
    gojson::ValuePtr root = gojson::DecodeAs(gojson::parse(machineBody, &perr), machineBody, gotypes::CMMachineData(), &perr);
    if (!root) {
        r.err = Error::New("failed to unmarshal CM get machine response body into machineData: " + perr);
        return r;
    }
    if (const Value* specs = cmResspecs(root.get()))
        for (const auto& s : specs->arr) {
            if (!cmSpecMatch(*s, specType, specModel)) continue;
            const std::string specUUID = s->get_string("spec_uuid");
            const long long count = s->get_int("device_count");
            if (const Value* devs = arrOf(s.get(), "devices"))
                for (const auto& d : devs->arr)
                    if (d->get_string("device_id") == deviceID) {
                        r.specUUID = specUUID;
                        r.deviceCount = count;
                        if (d->get_string("status") == "REMOVE_FAILED") r.err = Error::New(d->get_string("status_reason"));
                        return r;
                    }
            break;   // first matching spec only (:479)
        }
    return r;         // ("", 0, nil): nothing left to remove
}

// ---------------------------------------------------------------------------
// shared: Node -> Metal3Machine -> BareMetalHost -> machine uuid
// ---------------------------------------------------------------------------
Error FTIClientBase::machineIDFromAnnotations(const std::string& nodeName, bool useGivenNameInError, std::string* machineID) {
    K8sObject node;
    Error e = objects_->GetNode(nodeName, &node);
    if (!e.ok()) return e;
    const std::string machineInfo = node.annotations.count("machine.openshift.io/machine") ? node.annotations["machine.openshift.io/machine"] : "";
    const std::vector<std::string> mp = identity::Split(machineInfo, "/");
    if (mp.size() != 2)   // CM prints node.Name, FM the argument: the same string unless the store renames
        return Error::New("failed to get annotation 'machine.openshift.io/machine' from Node " +
                          (useGivenNameInError ? nodeName : node.name) + ", now is '" + machineInfo + "'");
    K8sObject machine;
    e = objects_->GetMetal3Machine(mp[0], mp[1], &machine);
    if (!e.ok()) return e;
    const std::string bmhInfo = machine.annotations.count("metal3.io/BareMetalHost") ? machine.annotations["metal3.io/BareMetalHost"] : "";
    const std::vector<std::string> bp = identity::Split(bmhInfo, "/");
    if (bp.size() != 2)
        return Error::New("failed to get annotation 'metal3.io/BareMetalHost' from Machine " + machine.name + ", now is '" + bmhInfo + "'");
    K8sObject bmh;
    e = objects_->GetBareMetalHost(bp[0], bp[1], &bmh);
    if (!e.ok()) return e;
    const std::string id = bmh.annotations.count("cluster-manager.cdi.io/machine") ? bmh.annotations["cluster-manager.cdi.io/machine"] : "";
    if (!bmh.has_annotations || id.empty())
        return Error::New("failed to get annotation 'cluster-manager.cdi.io/machine' from BareMetalHost " + bmh.name + ", now is '" + id + "'");
    *machineID = id;
    return Error::Nil();
}

// ---------------------------------------------------------------------------
// FM
// ---------------------------------------------------------------------------
namespace {
// url.Values{"tenant_uuid": {id}}.Encode() (fm/client.go:153-155 and twins)
std::string tenantQuery(const std::string& tenant) {
    static const char hex[] = "0123456789ABCDEF";
    std::string o = "tenant_uuid=";
    for (unsigned char c : tenant) {
        if ((c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '-' || c == '_' || c == '.' || c == '~') o.push_back((char)c);
        else if (c == ' ') o.push_back('+');
        else { o.push_back('%'); o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
    }
    return o;
}
}  // namespace

Error FMClient::getNodeMachineID(const std::string& nodeName, std::string* machineID) {
    if (!cfg_.clusterID.empty()) return machineIDFromAnnotations(nodeName, true, machineID);
    K8sObject node;                                                          // :451-463
    Error e = objects_->GetNode(nodeName, &node);
    if (!e.ok()) return e;
    const std::string prefix = "fsas-cdi://";
    if (node.provider_id.compare(0, prefix.size(), prefix) != 0)
        return Error::New("invalid format: expected 'fsas-cdi://machineUUID', now is '" + node.provider_id + "'");
    *machineID = node.provider_id.substr(prefix.size());
    return Error::Nil();
}

Error FMClient::getMachineInfo(const std::string& machineID, std::string* body) {
    Error e = token_->GetToken();
    if (!e.ok()) return e;
    HttpReply rep = send({"GET", "fabric_manager/api/v1/machines/" + machineID, tenantQuery(cfg_.tenantID), ""});
    if (!rep.transport_error.empty()) return Error::New(rep.transport_error);
    if (rep.status != 200) return FMErrorFromReply("get", rep.body);
    std::string perr;
    gojson::ValuePtr root = gojson::parse(rep.body, &perr);
    if (!gojson::decodesInto(root, rep.body, gotypes::FMGetMachineResponse(), &perr))
        return Error::New("failed to unmarshal FM get machine response body into machineData: " + perr);
    *body = rep.body;
    return Error::Nil();
}

Error FMClient::AddResource(const ComposableResource& instance, std::string* deviceID, std::string* CDIDeviceID) {
    std::string machineID;
    Error e = getNodeMachineID(instance.Spec.TargetNode, &machineID);
    if (!e.ok()) return e;
    e = token_->GetToken();
    if (!e.ok()) return e;
    HttpReply rep = send({"PATCH", "fabric_manager/api/v1/machines/" + machineID + "/update", tenantQuery(cfg_.tenantID),
                          FMScaleUpBody(cfg_.tenantID, machineID, instance.Spec.Type, instance.Spec.Model)});
    if (!rep.transport_error.empty()) return Error::New(rep.transport_error);
    if (rep.status != 200) return FMErrorFromReply("scaleup", rep.body);
    return controller::FMScaleUpResponseToIDs(rep.body, instance.Name, instance.Spec.Type, instance.Spec.Model, deviceID, CDIDeviceID);
}

Error FMClient::RemoveResource(ComposableResource& instance) {
    std::string machineID, machineBody;
    Error e = getNodeMachineID(instance.Spec.TargetNode, &machineID);
    if (!e.ok()) return e;
    e = getMachineInfo(machineID, &machineBody);
    if (!e.ok()) return e;
    std::string perr;
    gojson::ValuePtr root = gojson::DecodeAs(gojson::parse(machineBody, &perr), machineBody, gotypes::FMGetMachineResponse(), &perr);
    const Value* machines = arrOf(root ? root->get("data") : nullptr, "machines");
    if (!machines || machines->arr.empty())   // :231 indexes Machines[0] unguarded
        return Error::New("runtime error: index out of range [0] with length 0");
    bool exists = false;
    if (const Value* resources = arrOf(machines->arr[0].get(), "resources"))
        for (const auto& r : resources->arr)
            if (r->get_string("res_type") == instance.Spec.Type && r->get_string("res_uuid") == instance.Status.CDIDeviceID) {
                exists = true;
                break;
            }
    if (!exists) return Error::Nil();          // already gone: nothing to send (:238-241)
    e = token_->GetToken();
    if (!e.ok()) return e;
    HttpReply rep = send({"DELETE", "fabric_manager/api/v1/machines/" + machineID + "/update", tenantQuery(cfg_.tenantID),
                          FMScaleDownBody(cfg_.tenantID, machineID, instance.Spec.Type, instance.Status.CDIDeviceID)});
    if (!rep.transport_error.empty()) return Error::New(rep.transport_error);
    if (rep.status != 200 && rep.status != 204) return FMErrorFromReply("scaledown", rep.body);
    return Error::Nil();
}

Error FMClient::CheckResource(const ComposableResource& instance) {
    std::string machineID, machineBody;
    Error e = getNodeMachineID(instance.Spec.TargetNode, &machineID);
    if (!e.ok()) return e;
    e = getMachineInfo(machineID, &machineBody);
    if (!e.ok()) return e;
    return FMCheckResource(machineBody, instance.Spec.Type, instance.Spec.Model, instance.Status.DeviceID);
}

Error FMClient::GetResources(std::vector<DeviceInfo>* out) {
    std::vector<std::string> nodes;
    Error e = objects_->ListNodeNames(&nodes);
    if (!e.ok()) return e;
    out->clear();
    for (const auto& n : nodes) {                // per-node failures are logged and skipped (:373-383)
        std::string machineID, machineBody;
        if (!getNodeMachineID(n, &machineID).ok()) continue;
        if (!getMachineInfo(machineID, &machineBody).ok()) continue;
        FMGetResources(machineBody, n, machineID, out);
    }
    return Error::Nil();
}

// ---------------------------------------------------------------------------
// CM
// ---------------------------------------------------------------------------
Error CMClient::getMachineInfo(const std::string& machineID, std::string* body) {
    Error e = token_->GetToken();
    if (!e.ok()) return e;
    HttpReply rep = send({"GET", "cluster_manager/cluster_autoscaler/v3/tenants/" + cfg_.tenantID + "/clusters/" + cfg_.clusterID +
                                     "/machines/" + machineID, "", ""});
    if (!rep.transport_error.empty()) return Error::New(rep.transport_error);
    if (rep.status != 200) return CMErrorFromReply("get", rep.body);
    std::string perr;
    gojson::ValuePtr root = gojson::parse(rep.body, &perr);
    if (!gojson::decodesInto(root, rep.body, gotypes::CMMachineData(), &perr))
        return Error::New("failed to unmarshal CM get machine response body into machineData: " + perr);
    *body = rep.body;
    return Error::Nil();
}

Error CMClient::AddResource(const ComposableResource& instance, std::string* deviceID, std::string* CDIDeviceID) {
    std::string machineID, machineBody;
    Error e = machineIDFromAnnotations(instance.Spec.TargetNode, false, &machineID);
    if (!e.ok()) return e;
    e = getMachineInfo(machineID, &machineBody);
    if (!e.ok()) return e;
    std::vector<std::string> existing;
    e = objects_->ListComposableResourceDeviceIDs(&existing);
    if (!e.ok()) return e;
    controller::CMAddingResult r = controller::CMCheckAddingResources(machineBody, existing, instance.Spec.Type, instance.Spec.Model);
    if (!r.deviceID.empty()) {                   // an unused device is already there (:129-131)
        *deviceID = r.deviceID;
        *CDIDeviceID = r.CDIDeviceID;
        return r.err;
    }
    // the token is fetched after the request is built (:146-153); no request is sent without one
    e = token_->GetToken();
    if (!e.ok()) return e;
    HttpReply rep = send({"POST", "cluster_manager/cluster_autoscaler/v3/tenants/" + cfg_.tenantID + "/clusters/" + cfg_.clusterID +
                                      "/machines/" + machineID + "/actions/resize", "", CMScaleUpBody(r.specUUID, r.deviceCount + 1)});
    if (!rep.transport_error.empty()) return Error::New(rep.transport_error);
    if (rep.status != 200) return CMErrorFromReply("scaleup", rep.body);
    return Error::New(ErrWaitingDeviceAttaching);
}

Error CMClient::RemoveResource(ComposableResource& instance) {
    std::string machineID, machineBody;
    Error e = machineIDFromAnnotations(instance.Spec.TargetNode, false, &machineID);
    if (!e.ok()) return e;
    e = getMachineInfo(machineID, &machineBody);
    if (!e.ok()) return e;
    CMRemovingResult r = CMCheckRemovingResources(machineBody, instance.Spec.Type, instance.Spec.Model, instance.Status.DeviceID);
    if (!r.err.ok()) {                           // REMOVE_FAILED: record the reason, then try again (:199-207)
        instance.Status.Error = r.err.msg;
        e = objects_->UpdateStatus(instance);
        if (!e.ok()) return e;
    }
    if (r.specUUID.empty()) return Error::Nil();
    e = token_->GetToken();
    if (!e.ok()) return e;
    HttpReply rep = send({"POST", "cluster_manager/cluster_autoscaler/v3/tenants/" + cfg_.tenantID + "/clusters/" + cfg_.clusterID +
                                      "/machines/" + machineID + "/actions/resize", "",
                          CMScaleDownBody(r.specUUID, r.deviceCount - 1, instance.Status.DeviceID)});
    if (!rep.transport_error.empty()) return Error::New(rep.transport_error);
    if (rep.status != 200) return CMErrorFromReply("scaledown", rep.body);
    return Error::New(ErrWaitingDeviceDetaching);
}

Error CMClient::CheckResource(const ComposableResource& instance) {
    std::string machineID, machineBody;
    Error e = machineIDFromAnnotations(instance.Spec.TargetNode, false, &machineID);
    if (!e.ok()) return e;
    e = getMachineInfo(machineID, &machineBody);
    if (!e.ok()) return e;
    return CMCheckResource(machineBody, instance.Spec.Type, instance.Spec.Model, instance.Status.DeviceID);
}

Error CMClient::GetResources(std::vector<DeviceInfo>* out) {
    std::vector<std::string> nodes;
    Error e = objects_->ListNodeNames(&nodes);
    if (!e.ok()) return e;
    out->clear();
    for (const auto& n : nodes) {                // the CM flavour aborts on the first failure (:323-333)
        std::string machineID, machineBody;
        e = machineIDFromAnnotations(n, false, &machineID);
        if (!e.ok()) { out->clear(); return e; }
        e = getMachineInfo(machineID, &machineBody);
        if (!e.ok()) { out->clear(); return e; }
        CMGetResources(machineBody, n, machineID, out);
    }
    return Error::Nil();
}

// ---------------------------------------------------------------------------
// Sunfish
// ---------------------------------------------------------------------------
std::string SunfishBody(const std::string& name, long long count, const std::string& procType, const std::string& model) {
    gojson::Writer w;
    w.begin_object();
    w.field("Name", name);
    w.key("Processors").begin_object().key("Members").begin_array().begin_object();
    w.field("@Redfish.RequestCount", count);
    w.field("ProcessorType", procType);
    w.field("Model", model);
    w.end_object().end_array().end_object();
    w.end_object();
    return w.str();
}

Error SunfishClient::sendPatchRequest(const ComposableResource& instance, long long count) {
    // only the three models of :42-46 fill the ProcessorRequest; anything else sends its zero value (:108-115)
    const std::string& m = instance.Spec.Model;
    const bool known = m == "Tesla-V100-PCIE-16GB" || m == "NVIDIA-A100-PCIE-40GB" || m == "NVIDIA-A100-80GB-PCIe";
    HttpReply rep = send({"PATCH", "redfish/v1/Systems/System", "",
                          SunfishBody(instance.Spec.TargetNode, known ? count : 0, known ? "GPU" : "", known ? m : "")});
    if (!rep.transport_error.empty()) return Error::New(rep.transport_error);
    if (rep.status != 200 && rep.status != 204) return Error::New("http returned code " + std::to_string(rep.status));
    return Error::Nil();
}

Error SunfishClient::AddResource(const ComposableResource& instance, std::string* deviceID, std::string* CDIDeviceID) {
    deviceID->clear();
    CDIDeviceID->clear();
    return sendPatchRequest(instance, 1);
}

Error SunfishClient::RemoveResource(ComposableResource& instance) { return sendPatchRequest(instance, 0); }

}  // namespace fabric
}  // namespace cro
