// harness.cu — the reconcile step and the provider / node-side clients behind JSON entry points:
// cro_reconcile_attach, cro_fabric_list_devices, cro_local_node_op.  Everything the reference would have
// asked of the API server, the fabric or a kubelet arrives as data in the request (scripted fabric,
// scripted cluster) or is answered by the probe context; see include/croprobe.h for the request shapes.
#include <map>
#include <memory>
#include <set>
#include <string>

#include "../../include/croprobe.h"
#include "c_api_util.hpp"
#include "detach.hpp"
#include "fabric.hpp"
#include "gojson.hpp"
#include "gpus.hpp"
#include "identity.hpp"
#include "nodes.hpp"
#include "probe.hpp"
#include "provider.hpp"
#include "reconcile.hpp"

using namespace cro;
using namespace cro::capi;

extern "C" {

// ---- reconcile step ---------------------------------------------------------

namespace {

class JsonProvider : public controller::CdiProvider {
public:
    explicit JsonProvider(const gojson::Value* p) : p_(p) {}
    controller::Error AddResource(const controller::ComposableResource& inst, std::string* dev,
                                  std::string* cdi) override {
        if (!p_) return controller::Error::New("no provider configured");
        if (p_->get_bool("waiting")) return controller::Error::New(controller::ErrWaitingDeviceAttaching);
        const std::string err = p_->get_string("error");
        if (!err.empty()) return controller::Error::New(err);
        const gojson::Value* body = p_->get("fm_response_body");
        if (body && body->kind == gojson::Value::String)
            return controller::FMScaleUpResponseToIDs(body->str, inst.Name, inst.Spec.Type, inst.Spec.Model, dev, cdi);
        const gojson::Value* cm = p_->get("cm_machine_body");
        if (cm && cm->kind == gojson::Value::String) {
            // CM AddResource (fti/cm/client.go:107-182): unused device -> ids (+ error); none -> POST resize, wait
            std::vector<std::string> existing;
            const gojson::Value* ex = p_->get("existing_device_ids");
            if (ex && ex->kind == gojson::Value::Array)
                for (const auto& e : ex->arr)
                    if (e->kind == gojson::Value::String) existing.push_back(e->str);
            controller::CMAddingResult r = controller::CMCheckAddingResources(cm->str, existing, inst.Spec.Type, inst.Spec.Model);
            if (!r.deviceID.empty()) {
                *dev = r.deviceID;
                *cdi = r.CDIDeviceID;
                return r.err;
            }
            if (!r.err.ok()) return r.err;
            return controller::Error::New(controller::ErrWaitingDeviceAttaching);
        }
        *dev = p_->get_string("device_id");
        *cdi = p_->get_string("cdi_device_id");
        return controller::Error::Nil();
    }
    // Online state: the health decision over the GET-machine body (fabric codec), or a canned error
    controller::Error CheckResource(const controller::ComposableResource& inst) override {
        if (!p_) return controller::Error::Nil();
        const std::string canned = p_->get_string("check_resource_error");
        if (!canned.empty()) return controller::Error::New(canned);
        const gojson::Value* fm = p_->get("fm_machine_body");
        if (fm && fm->kind == gojson::Value::String)
            return fabric::FMCheckResource(fm->str, inst.Spec.Type, inst.Spec.Model, inst.Status.DeviceID);
        const gojson::Value* cm = p_->get("cm_check_body");
        if (cm && cm->kind == gojson::Value::String)
            return fabric::CMCheckResource(cm->str, inst.Spec.Type, inst.Spec.Model, inst.Status.DeviceID);
        return controller::Error::Nil();
    }
    controller::Error RemoveResource(controller::ComposableResource&) override {
        const gojson::Value* rm = p_ ? p_->get("remove") : nullptr;
        if (!rm) return controller::Error::Nil();
        if (rm->get_bool("waiting")) return controller::Error::New(controller::ErrWaitingDeviceDetaching);
        const std::string err = rm->get_string("error");
        return err.empty() ? controller::Error::Nil() : controller::Error::New(err);
    }

private:
    const gojson::Value* p_;
};

// ---- scripted fabric: the reference's httptest server + envtest objects, as data -------------
// "fabric": {"http": [{"method": "GET", "path_contains": "/machines/", "status": 404, "body": "..."}, ...],
//            "transport_error": "", "token_error": "",
//            "objects": {"nodes": {"worker-0": {"annotations": {..}, "provider_id": ""}},
//                        "metal3machines": {"ns/name": {"annotations": {..}}}, "baremetalhosts": {"ns/name": {..}},
//                        "composable_resource_device_ids": ["GPU-.."], "status_update_error": ""}}
class ScriptedTransport : public fabric::Transport {
public:
    explicit ScriptedTransport(const gojson::Value* f) : f_(f) {}
    fabric::HttpReply Do(const fabric::HttpRequest& req) override {
        fabric::HttpReply rep;
        if (f_) {
            const std::string terr = f_->get_string("transport_error");
            if (!terr.empty()) { rep.transport_error = terr; return rep; }
            const gojson::Value* rules = f_->get("http");
            if (rules && rules->kind == gojson::Value::Array)
                for (const auto& r : rules->arr) {
                    if (r->kind != gojson::Value::Object) continue;
                    const std::string m = r->get_string("method");
                    if (!m.empty() && m != req.method) continue;
                    const gojson::Value* exact = r->get("path");             // "path": whole path, like the
                    if (exact && exact->kind == gojson::Value::String) {     // reference's `switch r.URL.Path`
                        if (exact->str != req.path) continue;
                    } else if (req.path.find(r->get_string("path_contains")) == std::string::npos) {
                        continue;
                    }
                    rep.status = (int)r->get_int("status", 200);
                    rep.body = r->get_string("body");
                    return rep;
                }
        }
        rep.transport_error = req.method + " \"https://fabric/" + req.path + "\": no route in the scripted fabric";
        return rep;
    }

private:
    const gojson::Value* f_;
};

class JsonObjectStore : public fabric::ObjectStore {
public:
    explicit JsonObjectStore(const gojson::Value* f) : o_(f ? f->get("objects") : nullptr) {}
    std::vector<controller::ComposableResourceStatus>* updates = nullptr;

    controller::Error GetNode(const std::string& name, fabric::K8sObject* out) override {
        return get("nodes", "nodes", name, name, out);
    }
    controller::Error GetMetal3Machine(const std::string& ns, const std::string& name, fabric::K8sObject* out) override {
        return get("metal3machines", "metal3machines.infrastructure.cluster.x-k8s.io", ns + "/" + name, name, out);
    }
    controller::Error GetBareMetalHost(const std::string& ns, const std::string& name, fabric::K8sObject* out) override {
        return get("baremetalhosts", "baremetalhosts.metal3.io", ns + "/" + name, name, out);
    }
    controller::Error ListNodeNames(std::vector<std::string>* out) override {
        const gojson::Value* c = o_ ? o_->get("nodes") : nullptr;
        if (c && c->kind == gojson::Value::Object)
            for (const auto& kv : c->obj) out->push_back(kv.first);
        return controller::Error::Nil();
    }
    controller::Error ListComposableResourceDeviceIDs(std::vector<std::string>* out) override {
        const gojson::Value* c = o_ ? o_->get("composable_resource_device_ids") : nullptr;
        if (c && c->kind == gojson::Value::Array)
            for (const auto& e : c->arr)
                if (e->kind == gojson::Value::String) out->push_back(e->str);
        return controller::Error::Nil();
    }
    controller::Error UpdateStatus(const controller::ComposableResource& instance) override {
        const std::string e = o_ ? o_->get_string("status_update_error") : std::string();
        if (!e.empty()) return controller::Error::New(e);
        if (updates) updates->push_back(instance.Status);
        return controller::Error::Nil();
    }

private:
    controller::Error get(const char* coll, const char* resource, const std::string& key, const std::string& name,
                          fabric::K8sObject* out) {
        const gojson::Value* c = o_ ? o_->get(coll) : nullptr;
        const gojson::Value* v = (c && c->kind == gojson::Value::Object) ? c->get(key) : nullptr;
        if (!v || v->kind != gojson::Value::Object)   // apimachinery NewNotFound(gr, name).Error()
            return controller::Error::New(std::string(resource) + " \"" + name + "\" not found");
        out->name = name;
        out->provider_id = v->get_string("provider_id");
        const gojson::Value* a = v->get("annotations");
        if (a && a->kind == gojson::Value::Object) {
            out->has_annotations = true;
            for (const auto& kv : a->obj)
                if (kv.second->kind == gojson::Value::String) out->annotations[kv.first] = kv.second->str;
        }
        return controller::Error::Nil();
    }
    const gojson::Value* o_;
};

// "token_error": "<text GetToken returns>" (canned), or "token": {"secret_error","transport_error","status","body"}
// = what the id_manager answered, decoded by fabric::TokenFromReply (fti/token.go:96-175); "now" as elsewhere.
class JsonTokenSource : public fabric::TokenSource {
public:
    JsonTokenSource(const gojson::Value* f, const gojson::Value* in) : f_(f) {
        const gojson::Value* t = f ? f->get("token") : nullptr;
        if (t && t->kind == gojson::Value::Object) {
            fabric::TokenReply r;
            r.secret_error = t->get_string("secret_error");
            r.transport_error = t->get_string("transport_error");
            r.status = (int)t->get_int("status", 200);
            r.body = t->get_string("body");
            long long now = 0, ns = 0;
            std::string perr;
            nodes::ParseRFC3339(in ? in->get_string("now", "2025-01-01T00:00:00Z") : std::string("2025-01-01T00:00:00Z"), &now, &ns, &perr);
            real_.reset(new fabric::ReplyTokenSource(r, now));
        }
    }
    controller::Error GetToken() override {
        if (real_) return real_->GetToken();
        const std::string e = f_ ? f_->get_string("token_error") : std::string();
        return e.empty() ? controller::Error::Nil() : controller::Error::New(e);
    }
    int fetches() const { return real_ ? real_->fetches : 0; }

private:
    const gojson::Value* f_;
    std::unique_ptr<fabric::ReplyTokenSource> real_;
};

// ---- scripted cluster: the reference's envtest pods + gomonkey'd SPDY executor, as data -----------
// "cluster": {"cluster_policy": {"driver_enabled": true} | {} (spec.driver.enabled unset) | null (NotFound),
//             "pods": [{"namespace","name","node","labels":{..},"containers":[..]}],
//             "exec": [{"needle": {"escape": "<text, matched after net/url.QueryEscape>"} | {"literal": "<raw query text>"} | null,
//                       "stdout","stderr","exec_err"}, ...]}      first matching rule answers (`strings.Contains(url.RawQuery, needle)`)
// ResourceSlices come from the request's "resource_slices".
class JsonKube : public gpus::Kube {
public:
    JsonKube(const gojson::Value* cluster, const gojson::Value* in) : c_(cluster), in_(in) {}
    controller::Error GetClusterPolicy(bool* found, bool* set, bool* enabled) override {
        const gojson::Value* cp = c_->get("cluster_policy");
        *found = cp && cp->kind == gojson::Value::Object;
        const gojson::Value* en = *found ? cp->get("driver_enabled") : nullptr;
        *set = en && en->kind == gojson::Value::Bool;
        *enabled = *set && en->b;
        const std::string err = c_->get_string("cluster_policy_error");
        return err.empty() ? controller::Error::Nil() : controller::Error::New(err);
    }
    controller::Error ListPods(std::vector<gpus::Pod>* out) override {
        const gojson::Value* pods = c_->get("pods");
        if (pods && pods->kind == gojson::Value::Array)
            for (const auto& p : pods->arr) {
                if (p->kind != gojson::Value::Object) continue;
                gpus::Pod pod;
                pod.ns = p->get_string("namespace");
                pod.name = p->get_string("name");
                pod.node = p->get_string("node");
                if (const gojson::Value* l = p->get("labels"))
                    if (l->kind == gojson::Value::Object)
                        for (const auto& kv : l->obj)
                            if (kv.second->kind == gojson::Value::String) pod.labels[kv.first] = kv.second->str;
                if (const gojson::Value* cs = p->get("containers"))
                    if (cs->kind == gojson::Value::Array)
                        for (const auto& cn : cs->arr)
                            if (cn->kind == gojson::Value::String) pod.containers.push_back(cn->str);
                out->push_back(pod);
            }
        return controller::Error::Nil();
    }
    controller::Error ListResourceSliceUUIDs(std::vector<std::string>* out) override {
        const gojson::Value* slices = in_->get("resource_slices");
        if (slices && slices->kind == gojson::Value::Array)
            for (const auto& s : slices->arr) {
                const gojson::Value* devs = s->get("devices");
                if (!devs || devs->kind != gojson::Value::Array) continue;
                for (const auto& d : devs->arr) {
                    const gojson::Value* attrs = d->get("attributes");
                    if (attrs && attrs->kind == gojson::Value::Object) {
                        const std::string u = attrs->get_string("uuid");
                        if (attrs->get("uuid")) out->push_back(u);
                    }
                }
            }
        return controller::Error::Nil();
    }

    // DeviceTaintRules: "taints": ["<name>", ..] exist already; "taint_get_error" / "taint_create_error" /
    // "taint_delete_error" make the corresponding API call fail; every create / delete is logged.
    controller::Error ListResourceSliceDevices(std::vector<SliceDevice>* out) override {
        const gojson::Value* slices = in_->get("resource_slices");
        if (slices && slices->kind == gojson::Value::Array)
            for (const auto& s : slices->arr) {
                const gojson::Value* pool = s->get("pool");
                const gojson::Value* devs = s->get("devices");
                if (!devs || devs->kind != gojson::Value::Array) continue;
                for (const auto& d : devs->arr) {
                    const gojson::Value* attrs = d->get("attributes");
                    if (!attrs || attrs->kind != gojson::Value::Object || !attrs->get("uuid")) continue;
                    out->push_back({s->get_string("driver"), pool ? pool->get_string("name") : std::string(), d->get_string("name"),
                                    attrs->get_string("uuid")});
                }
            }
        return controller::Error::Nil();
    }
    controller::Error GetDeviceTaintRule(const std::string& name, bool* found) override {
        *found = created_.count(name) > 0;
        const gojson::Value* t = c_->get("taints");
        if (t && t->kind == gojson::Value::Array)
            for (const auto& e : t->arr)
                if (e->kind == gojson::Value::String && e->str == name && !deleted_.count(name)) *found = true;
        return fail("taint_get_error");
    }
    controller::Error CreateDeviceTaintRule(const TaintRule& r) override {
        controller::Error e = fail("taint_create_error");
        if (!e.ok()) return e;
        created_.insert(r.name);
        taint_ops.push_back("create " + r.name + " driver=" + r.driver + " pool=" + r.pool + " device=" + r.device + " " + r.key + "=" +
                            r.value + ":" + r.effect);
        return e;
    }
    controller::Error DeleteDeviceTaintRule(const std::string& name) override {
        controller::Error e = fail("taint_delete_error");
        if (!e.ok()) return e;
        created_.erase(name);
        deleted_.insert(name);
        taint_ops.push_back("delete " + name);
        return e;
    }
    std::vector<std::string> taint_ops;

private:
    controller::Error fail(const char* key) {
        const std::string e = c_->get_string(key);
        return e.empty() ? controller::Error::Nil() : controller::Error::New(e);
    }
    std::set<std::string> created_, deleted_;
    const gojson::Value* c_;
    const gojson::Value* in_;
};

class ScriptedExec : public gpus::Exec {
public:
    explicit ScriptedExec(const gojson::Value* cluster) : c_(cluster) {}
    struct Entry { std::string pod, container, query; std::vector<std::string> argv; int kind; bool detached; };
    std::vector<Entry> log;
    int slept = 0;
    void Sleep(int s) override { slept += s; }
    gpus::ExecResult Run(const gpus::Pod& pod, const std::string& container, const gpus::ExecRequest& req) override {
        const std::vector<std::string> argv = req.kind == gpus::ExecRequest::Command ? req.argv : gpus::ScanAsCommand(req);
        const std::string query = gpus::ExecRawQuery(argv, container);
        log.push_back({pod.ns + "/" + pod.name, container, query, argv, (int)req.kind, req.detached});
        gpus::ExecResult r;
        const gojson::Value* rules = c_->get("exec");
        if (rules && rules->kind == gojson::Value::Array)
            for (const auto& rule : rules->arr) {
                if (rule->kind != gojson::Value::Object) continue;
                const gojson::Value* n = rule->get("needle");
                if (n && n->kind == gojson::Value::Object) {
                    const gojson::Value* esc = n->get("escape");
                    const gojson::Value* lit = n->get("literal");
                    std::string needle;
                    if (esc && esc->kind == gojson::Value::String) {
                        // net/url.QueryEscape of the text, the same way ExecRawQuery escapes each argument
                        needle = gpus::ExecRawQuery({esc->str}, "").substr(8);       // strip "command="
                        needle = needle.substr(0, needle.find("&container="));
                    } else if (lit && lit->kind == gojson::Value::String) {
                        needle = lit->str;
                    }
                    if (query.find(needle) == std::string::npos) continue;
                }
                r.std_out = rule->get_string("stdout");
                r.std_err = rule->get_string("stderr");
                const gojson::Value* ee = rule->get("exec_err");
                if (ee && ee->kind == gojson::Value::String) { r.failed = true; r.exec_err = ee->str; }
                return r;
            }
        r.failed = true;
        r.exec_err = "no exec rule matches " + query;
        return r;
    }

private:
    const gojson::Value* c_;
};

// NewComposableResourceAdapter (composableresource_adapter.go:39-72) over an env map: which
// provider flavour, or the error the reconcile surfaces.  "" kind + nil error never happens.
controller::Error SelectAdapter(const gojson::Value* env, std::string* kind) {
    const std::string drt = env->get_string("DEVICE_RESOURCE_TYPE");
    if (drt != "DEVICE_PLUGIN" && drt != "DRA")
        return controller::Error::New("the env variable DEVICE_RESOURCE_TYPE has an invalid value: '" + drt + "'");
    const std::string provider = env->get_string("CDI_PROVIDER_TYPE");
    if (provider == "SUNFISH") { *kind = "sunfish"; return controller::Error::Nil(); }
    if (provider == "FTI_CDI") {
        if (env->get_string("FTI_CDI_CLUSTER_ID").empty() && drt == "DEVICE_PLUGIN")
            return controller::Error::New("The cluster in RKE2 does not support DEVICE_PLUGIN, please use DRA");
        const std::string api = env->get_string("FTI_CDI_API_TYPE");
        if (api == "CM") { *kind = "cm"; return controller::Error::Nil(); }
        if (api == "FM") { *kind = "fm"; return controller::Error::Nil(); }
        return controller::Error::New("the env variable FTI_CDI_API_TYPE has an invalid value: '" + api + "'");
    }
    return controller::Error::New("the env variable CDI_PROVIDER_TYPE has an invalid value: '" + provider + "'");
}

class ProbeNodeOps : public controller::NodeOps {
public:
    ProbeNodeOps(cro_ctx* ctx, const gojson::Value* in) : ctx_(ctx), in_(in) {}
    bool probed = false;
    cro_probe_result probe_result{};

    controller::Error CheckNoGPULoads(const std::string&) override { return controller::Error::Nil(); }
    controller::Error RestartDaemonset(const std::string& ns, const std::string& name) override {
        const gojson::Value* errs = in_->get("daemonset_errors");
        if (errs && errs->kind == gojson::Value::Object) {
            const std::string e = errs->get_string(ns + "/" + name);
            if (!e.empty()) return controller::Error::New(e);
        }
        // "daemonsets": {"ns/name": {"desired":1,"ready":1,"current":1,"unavailable":0,"misscheduled":0,
        //                            "restarted_at":"2025-01-01T00:00:00Z"}}, "now": "<RFC3339>":
        // the restart rule of internal/utils/nodes.go:35-76 decides; absent objects keep the old "restart ok".
        const gojson::Value* sets = in_->get("daemonsets");
        const gojson::Value* ds = (sets && sets->kind == gojson::Value::Object) ? sets->get(ns + "/" + name) : nullptr;
        if (sets && sets->kind == gojson::Value::Object && !ds)     // client.Get NotFound
            return controller::Error::New("daemonsets.apps \"" + name + "\" not found");
        if (!ds || ds->kind != gojson::Value::Object) return controller::Error::Nil();
        nodes::DaemonSetView v;
        v.DesiredNumberScheduled = ds->get_int("desired");
        v.NumberReady = ds->get_int("ready");
        v.CurrentNumberScheduled = ds->get_int("current");
        v.NumberUnavailable = ds->get_int("unavailable");
        v.NumberMisscheduled = ds->get_int("misscheduled");
        if (const gojson::Value* ra = ds->get("restarted_at"))
            if (ra->kind == gojson::Value::String) { v.hasRestartedAt = true; v.restartedAt = ra->str; }
        long long now = 0, nowNs = 0;
        std::string perr;
        const std::string nowText = in_->get_string("now", "2025-01-01T00:00:00Z");
        if (!nodes::ParseRFC3339(nowText, &now, &nowNs, &perr)) return controller::Error::New("bad \"now\": " + perr);
        nodes::Restart what;
        controller::Error e = nodes::RestartDaemonsetDecision(ns, name, v, now, nowNs, &what);
        if (e.ok() && what == nodes::Restart::Restarted) restarted.push_back(ns + "/" + name + "@" + nodes::FormatRFC3339UTC(now));
        return e;
    }
    std::vector<std::string> restarted;   // "ns/name@<restartedAt stamp>" for every Update the rule issued
    controller::Error RunNvidiaSmi(const std::string& node) override {
        std::vector<std::string> uuids;
        return enumerate(node, &uuids);
    }
    // ---- detach side: the text rules of csrc/detach.cpp over injected command output ----
    static bool io(const gojson::Value* v, std::string* so, std::string* se, std::string* ee_s, const char** ee) {
        if (!v || v->kind != gojson::Value::Object) return false;
        *so = v->get_string("stdout");
        *se = v->get_string("stderr");
        const gojson::Value* e = v->get("exec_err");
        *ee = nullptr;
        if (e && e->kind == gojson::Value::String) { *ee_s = e->str; *ee = ee_s->c_str(); }
        return true;
    }
    controller::Error CheckNoGPULoadsFor(const std::string& node, const std::string* uuid) override {
        std::string so, se, ee_s;
        const char* ee;
        const gojson::Value* lc = in_->get("load_check");
        if (!io(lc, &so, &se, &ee_s, &ee)) return controller::Error::Nil();
        return detach::CheckNoGPULoadsFromOutput(so, se, ee, lc->get_string("pod_name", "nvidia-driver-daemonset-test"), node, uuid,
                                                 lc->get_bool("driver_enabled", true));
    }
    controller::Error CreateDeviceTaint(const controller::ComposableResource&) override {
        const std::string e = in_->get_string("create_taint_error");
        return e.empty() ? controller::Error::Nil() : controller::Error::New(e);
    }
    controller::Error DeleteDeviceTaint(const controller::ComposableResource&) override {
        const std::string e = in_->get_string("delete_taint_error");
        return e.empty() ? controller::Error::Nil() : controller::Error::New(e);
    }
    controller::Error DrainGPU(const std::string&, const std::string&, const std::string&) override {
        const gojson::Value* dr = in_->get("drain");
        if (!dr) return controller::Error::Nil();
        const std::string canned = dr->get_string("error");
        if (!canned.empty()) return controller::Error::New(canned);
        std::string so, se, ee_s;
        const char* ee;
        if (io(dr->get("fd_scan"), &so, &se, &ee_s, &ee))
            return detach::CheckDeviceFileScanResult(so, se, ee, dr->get_bool("rke2"));
        return controller::Error::Nil();
    }
    controller::Error CheckGPUVisible(const std::string& type, const controller::ComposableResource& r,
                                      bool* visible) override {
        *visible = false;
        bool listed = false;
        wanted_ = r.Status.DeviceID;
        // while detaching, the cluster is looked at AFTER the fabric removed the device
        const bool after = r.Status.State == "Detaching";
        const gojson::Value* slices = in_->get(after && in_->get("resource_slices_after_remove") ? "resource_slices_after_remove" : "resource_slices");
        if (type == "DRA" && slices && slices->kind == gojson::Value::Array) {
            // internal/utils/gpus.go:55-71
            for (const auto& rs : slices->arr) {
                const gojson::Value* devs = rs->get("devices");
                if (!devs || devs->kind != gojson::Value::Array) continue;
                for (const auto& d : devs->arr) {
                    const gojson::Value* attrs = d->get("attributes");
                    if (attrs && attrs->get_string("uuid") == r.Status.DeviceID && attrs->get("uuid")) listed = true;
                }
            }
        } else {
            std::vector<std::string> uuids;
            controller::Error e = enumerate(r.Spec.TargetNode, &uuids, after && in_->get("enumeration_after_remove") ? "enumeration_after_remove" : "enumeration");
            if (!e.ok()) return e;
            for (const std::string& u : uuids)
                if (u == r.Status.DeviceID) listed = true;   // gpus.go:78-82
        }
        if (!listed) return controller::Error::Nil();
        if (in_->get_bool("probe") && ctx_ && !after) {
            // the strong check: the device must also deliver its HBM pattern
            // by UUID: a device this context holds is probed in process, one that reached the node after
            // cro_probe_init through the helper process; CRO_ERR_NO_DEVICE = not on the node (any more)
            int rc = ctx_probe_uuid(ctx_, r.Status.DeviceID.c_str(), &probe_result);
            if (rc == CRO_ERR_NO_DEVICE) return controller::Error::Nil();
            probed = true;
            if (rc != CRO_OK) {
                char msg[512] = {0};
                cro_last_error(ctx_, msg, sizeof msg);
                return controller::Error::New(std::string("cuda probe failed: ") + cro_strerror(rc) +
                                              (msg[0] ? std::string(": ") + msg : std::string()));
            }
        }
        *visible = true;
        return controller::Error::Nil();
    }

private:
    controller::Error enumerate(const std::string& node, std::vector<std::string>* uuids, const char* key = "enumeration") {
        if (in_->get_bool("driver_pod_missing"))   // gpus.go:835
            return controller::Error::New(
                "no Pod with label 'app.kubernetes.io/component=nvidia-driver' found on node " + node);
        std::string out, err;
        const char* exec_err = nullptr;
        std::string exec_err_s;
        const gojson::Value* en = in_->get(key);
        if (en && en->kind == gojson::Value::Object) {
            out = en->get_string("stdout");
            err = en->get_string("stderr");
            const gojson::Value* ee = en->get("exec_err");
            if (ee && ee->kind == gojson::Value::String) { exec_err_s = ee->str; exec_err = exec_err_s.c_str(); }
        } else if (ctx_) {
            // a FRESH look at the node on every reconcile, like the reference's exec of nvidia-smi (gpus.go:886)
            std::vector<cro_dev_info> infos;
            ctx_inventory(ctx_, &infos);
            if (!wanted_.empty() && !identity::CheckGPUVisible(infos.data(), (int)infos.size(), wanted_))
                ctx_inventory(ctx_, &infos, true);      // the fabric named a device the quick look does not show: look properly
            identity::EmitCsv(infos.data(), (int)infos.size(), "gpu_uuid", &out, nullptr);
        } else {
            return controller::Error::New("no probe context and no enumeration text");
        }
        identity::GpuInfoResult r = identity::getGPUInfoFromNvidiaSmiOutput(out, err, exec_err, "gpu_uuid");
        if (r.code != CRO_OK) return controller::Error::New(r.error);
        for (const auto& g : r.infos) {
            auto it = g.find("gpu_uuid");
            if (it != g.end()) uuids->push_back(it->second);
        }
        return controller::Error::Nil();
    }
    cro_ctx* ctx_;
    const gojson::Value* in_;
    std::string wanted_;      // Status.DeviceID of the resource being reconciled, once known
};

}  // namespace

int cro_reconcile_attach(cro_ctx* ctx, const char* in_json, char* buf, size_t cap, size_t* len) try {
    if (!in_json) return CRO_ERR_INVALID_ARG;
    std::string perr;
    gojson::ValuePtr in = gojson::parse(in_json, &perr);
    if (!in || in->kind != gojson::Value::Object) {
        copy_out("bad reconcile request: " + perr, buf, cap, len);
        return CRO_ERR_PARSE;
    }
    controller::ComposableResource res;
    res.Name = in->get_string("name", "test-composable-resource");
    res.DeletionTimestampSet = in->get_bool("deleting");
    if (const gojson::Value* sp = in->get("spec")) {
        res.Spec.Type = sp->get_string("type");
        res.Spec.Model = sp->get_string("model");
        res.Spec.TargetNode = sp->get_string("target_node");
        res.Spec.ForceDetach = sp->get_bool("force_detach");
    }
    if (const gojson::Value* st = in->get("status")) {
        res.Status.State = st->get_string("state");
        res.Status.Error = st->get_string("error");
        res.Status.DeviceID = st->get_string("device_id");
        res.Status.CDIDeviceID = st->get_string("cdi_device_id");
    }
    if (const gojson::Value* lb = in->get("labels"))
        if (lb->kind == gojson::Value::Object)
            for (const auto& kv : lb->obj)
                if (kv.second->kind == gojson::Value::String) res.Labels[kv.first] = kv.second->str;
    std::string type = in->get_string("device_resource_type", "DEVICE_PLUGIN");

    // With an "env" object the adapter is chosen the way the operator does it and the provider is
    // the real FM / CM client over the scripted fabric; otherwise the canned JsonProvider.
    const gojson::Value* env = in->get("env");
    const gojson::Value* fab = in->get("fabric");
    ScriptedTransport transport(fab);
    JsonObjectStore store(fab);
    JsonTokenSource tokens(fab, in.get());
    std::unique_ptr<fabric::FTIClientBase> fti;
    controller::Error adapterErr;
    if (env && env->kind == gojson::Value::Object) {
        type = env->get_string("DEVICE_RESOURCE_TYPE");
        std::string kind;
        adapterErr = SelectAdapter(env, &kind);
        fabric::ClientConfig cfg{env->get_string("FTI_CDI_TENANT_ID"), env->get_string("FTI_CDI_CLUSTER_ID")};
        if (kind == "fm") fti.reset(new fabric::FMClient(cfg, &transport, &store, &tokens));
        else if (kind == "cm") fti.reset(new fabric::CMClient(cfg, &transport, &store, &tokens));
        else if (kind == "sunfish") fti.reset(new fabric::SunfishClient(&transport));
        else if (adapterErr.ok()) adapterErr = controller::Error::New("provider kind '" + kind + "' is not scripted in this harness");
    }

    JsonProvider canned(in->get("provider"));
    controller::CdiProvider* provider = fti ? static_cast<controller::CdiProvider*>(fti.get()) : &canned;
    ProbeNodeOps node(ctx, in.get());
    // With a "cluster" object the node side is csrc/gpus.cpp (internal/utils/gpus.go restated) over the
    // scripted pods / pod-exec; DaemonSet restarts and taints stay with the canned NodeOps above.
    const gojson::Value* cluster = in->get("cluster");
    const bool scripted = cluster && cluster->kind == gojson::Value::Object;
    JsonKube kube(scripted ? cluster : in.get(), in.get());
    ScriptedExec pod_exec(scripted ? cluster : in.get());
    struct ClusterNodeOps : gpus::GpuNodeOps {
        ClusterNodeOps(gpus::Kube* k, gpus::Exec* e, ProbeNodeOps* canned) : gpus::GpuNodeOps(k, e), canned_(canned) {}
        controller::Error RestartDaemonset(const std::string& ns, const std::string& name) override { return canned_->RestartDaemonset(ns, name); }
        ProbeNodeOps* canned_;
    } cluster_ops(&kube, &pod_exec, &node);
    controller::NodeOps* node_ops = scripted ? static_cast<controller::NodeOps*>(&cluster_ops) : &node;
    controller::ComposableResourceReconciler rec(provider, node_ops);
    store.updates = &rec.statusUpdates;
    // "status_update_failures": {"after": N, "error": "..."} — Status().Update succeeds N times, then answers the error
    // (the API server refusing a write: conflict, webhook, etcd); the handler stops where the reference stops
    long long writes_ok = -1, writes_seen = 0;
    std::string write_error;
    if (const gojson::Value* wf = in->get("status_update_failures")) {
        writes_ok = wf->get_int("after");
        write_error = wf->get_string("error", "Operation cannot be fulfilled");
    }
    rec.writer = [&](const controller::ComposableResource&) {
        ++writes_seen;
        return (writes_ok >= 0 && writes_seen > writes_ok) ? controller::Error::New(write_error) : controller::Error::Nil();
    };
    controller::Result result;
    controller::Error err;
    if (!adapterErr.ok()) {
        err = rec.requeueOnErr(&res, adapterErr);          // composableresource_controller.go:91-94
    } else if (type != "DEVICE_PLUGIN" && type != "DRA") {
        // composableresource_adapter.go:42-45
        err = rec.requeueOnErr(&res, controller::Error::New(
                                         "the env variable DEVICE_RESOURCE_TYPE has an invalid value: '" + type + "'"));
    } else if (res.Status.State.empty()) {
        err = rec.handleNoneState(&res, &result);
    } else if (res.Status.State == "Attaching") {
        err = rec.handleAttachingState(&res, type, &result);
    } else if (res.Status.State == "Online") {
        err = rec.handleOnlineState(&res, &result);
    } else if (res.Status.State == "Detaching") {
        err = rec.handleDetachingState(&res, type, &result);
    }

    gojson::Writer w;
    w.begin_object();
    w.key("status").raw(res.Status.MarshalJSON());
    w.field("requeue_after_s", result.RequeueAfterSeconds);
    w.field("delete_requested", res.DeleteRequested);
    w.field("error", err.ok() ? std::string() : err.msg);
    w.key("status_updates").begin_array();
    for (const auto& s : rec.statusUpdates) w.raw(s.MarshalJSON());
    w.end_array();
    w.field("failed_status_updates", (long long)rec.failedUpdates);
    if (node.probed) w.key("probe").string_map(probe_annotations(node.probe_result));
    if (scripted) {     // every pod-exec the step issued, in order: pod, container, the URL query the mocks match on, argv
        w.key("exec_log").begin_array();
        for (const auto& x : pod_exec.log) {
            w.begin_object();
            w.field("pod", x.pod).field("container", x.container).field("query", x.query);
            w.key("argv").begin_array();
            for (const auto& a : x.argv) w.value(a);
            w.end_array();
            w.field("kind", x.kind == 0 ? std::string("command") : x.kind == 1 ? std::string("fd_scan") : x.kind == 2 ? std::string("proc_scan") : std::string("cmdline_scan"));
            w.field("detached", x.detached);
            w.end_object();
        }
        w.end_array();
        w.field("slept_s", pod_exec.slept);
        w.key("taint_ops").begin_array();
        for (const auto& op : kube.taint_ops) w.value(op);
        w.end_array();
    }
    if (!node.restarted.empty()) {
        w.key("daemonset_restarts").begin_array();
        for (const auto& r : node.restarted) w.value(r);
        w.end_array();
    }
    if (fti) {
        w.field("token_fetches", tokens.fetches());
        w.key("fabric_requests").begin_array();
        for (const auto& r : fti->requests) {
            w.begin_object();
            w.field("method", r.method).field("path", r.path).field("query", r.query).field("body", r.body);
            w.end_object();
        }
        w.end_array();
    }
    w.end_object();
    return copy_out(w.str(), buf, cap, len);
} CRO_API_CATCH

int cro_fabric_list_devices(const char* request_json, char* buf, size_t cap, size_t* len) try {
    if (!request_json) return CRO_ERR_INVALID_ARG;
    std::string perr;
    gojson::ValuePtr in = gojson::parse(request_json, &perr);
    if (!in || in->kind != gojson::Value::Object) {
        copy_out("bad request: " + perr, buf, cap, len);
        return CRO_ERR_PARSE;
    }
    const gojson::Value* env = in->get("env");
    const gojson::Value* fab = in->get("fabric");
    if (!env || env->kind != gojson::Value::Object) return CRO_ERR_INVALID_ARG;
    ScriptedTransport transport(fab);
    JsonObjectStore store(fab);
    JsonTokenSource tokens(fab, in.get());
    std::string kind;
    controller::Error e = SelectAdapter(env, &kind);
    std::unique_ptr<fabric::FTIClientBase> fti;
    fabric::ClientConfig cfg{env->get_string("FTI_CDI_TENANT_ID"), env->get_string("FTI_CDI_CLUSTER_ID")};
    if (kind == "fm") fti.reset(new fabric::FMClient(cfg, &transport, &store, &tokens));
    else if (kind == "cm") fti.reset(new fabric::CMClient(cfg, &transport, &store, &tokens));
    else if (kind == "sunfish") fti.reset(new fabric::SunfishClient(&transport));
    else if (e.ok()) e = controller::Error::New("provider kind '" + kind + "' is not scripted in this harness");
    std::vector<fabric::DeviceInfo> devs;
    if (fti) e = fti->GetResources(&devs);
    gojson::Writer w;
    w.begin_object();
    w.key("devices").raw(fabric::DeviceInfosToJson(devs));
    w.field("error", e.ok() ? std::string() : e.msg);
    w.field("token_fetches", tokens.fetches());
    w.key("fabric_requests").begin_array();
    if (fti)
        for (const auto& r : fti->requests) {
            w.begin_object();
            w.field("method", r.method).field("path", r.path).field("query", r.query).field("body", r.body);
            w.end_object();
        }
    w.end_array();
    w.end_object();
    return copy_out(w.str(), buf, cap, len);
} CRO_API_CATCH

int cro_token_from_reply(const char* reply_json, char* buf, size_t cap, size_t* len) try {
    if (!reply_json) return CRO_ERR_INVALID_ARG;
    std::string perr;
    gojson::ValuePtr in = gojson::parse(reply_json, &perr);
    if (!in || in->kind != gojson::Value::Object) {
        copy_out("bad request: " + perr, buf, cap, len);
        return CRO_ERR_PARSE;
    }
    fabric::TokenReply r;
    r.secret_error = in->get_string("secret_error");
    r.transport_error = in->get_string("transport_error");
    r.status = (int)in->get_int("status", 200);
    r.body = in->get_string("body");
    long long exp = 0;
    controller::Error e = fabric::TokenFromReply(r, &exp);
    gojson::Writer w;
    w.begin_object();
    w.field("error", e.ok() ? std::string() : e.msg);
    w.field("expiry", e.ok() ? exp : 0LL);
    w.end_object();
    return copy_out(w.str(), buf, cap, len);
} CRO_API_CATCH

// ---- node-side operations on the node itself ----------------------------------------

int cro_scan_cmdline_for(const char* proc_root, const char* needle, int* found) try {
    if (!needle || !found) return CRO_ERR_INVALID_ARG;
    *found = detach::ScanCmdlineFor(S(proc_root), needle).empty() ? 0 : 1;
    return CRO_OK;
} CRO_API_CATCH

int cro_local_exec(const char* request_json, char* buf, size_t cap, size_t* len) try {
    if (!request_json) return CRO_ERR_INVALID_ARG;
    std::string perr;
    gojson::ValuePtr in = gojson::parse(request_json, &perr);
    if (!in || in->kind != gojson::Value::Object) {
        copy_out("bad request: " + perr, buf, cap, len);
        return CRO_ERR_PARSE;
    }
    gpus::LocalExec::Options o;
    o.allow_mutation = in->get_bool("allow_mutation");
    if (in->get("exec_deadline_ms")) o.exec_deadline_ms = (int)in->get_int("exec_deadline_ms");
    if (in->get("native_nvml")) o.native_nvml = in->get_bool("native_nvml");
    o.nvml_lib = in->get_string("nvml_lib");
    gpus::LocalExec exec(o);
    gpus::ExecRequest req;
    req.kind = gpus::ExecRequest::Command;
    if (const gojson::Value* av = in->get("argv"))
        for (const auto& a : av->arr)
            if (a->kind == gojson::Value::String) req.argv.push_back(a->str);
    gpus::Pod pod;
    const gpus::ExecResult r = exec.Run(pod, "", req);
    gojson::Writer w;
    w.begin_object();
    w.field("how", exec.log.empty() ? std::string() : exec.log.back().how);
    w.field("failed", r.failed).field("exec_err", r.exec_err).field("stdout", r.std_out).field("stderr", r.std_err);
    w.end_object();
    return copy_out(w.take(), buf, cap, len);
} CRO_API_CATCH

int cro_local_node_op(cro_ctx* ctx, const char* request_json, char* buf, size_t cap, size_t* len) try {
    if (!request_json) return CRO_ERR_INVALID_ARG;
    std::string perr;
    gojson::ValuePtr in = gojson::parse(request_json, &perr);
    if (!in || in->kind != gojson::Value::Object) {
        copy_out("bad request: " + perr, buf, cap, len);
        return CRO_ERR_PARSE;
    }
    std::vector<cro_dev_info> devs;
    if (ctx) ctx_inventory(ctx, &devs);      // never the init-time snapshot: a drained GPU must stop being listed
    gpus::LocalExec::Options o;
    o.proc_root = in->get_string("proc_root");
    o.allow_mutation = in->get_bool("allow_mutation");
    if (in->get("exec_deadline_ms")) o.exec_deadline_ms = (int)in->get_int("exec_deadline_ms");
    if (ctx) { o.devs = devs.data(); o.n_devs = (int)devs.size(); }
    if (in->get("native_nvml")) o.native_nvml = in->get_bool("native_nvml");
    o.nvml_lib = in->get_string("nvml_lib");
    gpus::LocalExec exec(o);
    const std::string node = in->get_string("node", "local");
    gpus::LocalKube kube(node, in->get_bool("driver_container", true));
    gpus::GpuNodeOps ops(&kube, &exec);
    const std::string op = in->get_string("op");
    const std::string uuid = in->get_string("device_id");
    const std::string type = in->get_string("device_resource_type", "DEVICE_PLUGIN");
    controller::Error e;
    bool visible = false;
    if (op == "check_no_gpu_loads") e = ops.CheckNoGPULoadsFor(node, uuid.empty() ? nullptr : &uuid);
    else if (op == "run_nvidia_smi") e = ops.RunNvidiaSmi(node);
    else if (op == "check_gpu_visible") {
        controller::ComposableResource r;
        r.Spec.TargetNode = node;
        r.Status.DeviceID = uuid;
        e = ops.CheckGPUVisible(type, r, &visible);
    } else if (op == "drain") e = ops.DrainGPU(node, uuid, type);
    else return CRO_ERR_INVALID_ARG;
    gojson::Writer w;
    w.begin_object();
    w.field("error", e.ok() ? std::string() : e.msg);
    w.field("visible", visible);
    w.key("exec_log").begin_array();
    for (const auto& x : exec.log) {
        w.begin_object();
        w.field("kind", x.kind == 0 ? std::string("command") : x.kind == 1 ? std::string("fd_scan") : x.kind == 2 ? std::string("proc_scan") : std::string("cmdline_scan"));
        w.key("argv").begin_array();
        for (const auto& a : x.argv) w.value(a);
        w.end_array();
        w.field("how", x.how).field("failed", x.failed);
        w.end_object();
    }
    w.end_array();
    w.end_object();
    return copy_out(w.str(), buf, cap, len);
} CRO_API_CATCH

}  // extern "C"
