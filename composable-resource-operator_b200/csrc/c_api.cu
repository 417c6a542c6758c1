// c_api.cu — extern "C" surface of libcroprobe (include/croprobe.h).
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>

#include "../../include/croprobe.h"
#include "c_api_util.hpp"
#include "gojson.hpp"
#include "identity.hpp"
#include "probe.hpp"
#include "reconcile.hpp"
#include "cluster.hpp"
#include "detach.hpp"
#include "fabric.hpp"
#include "provider.hpp"
#include "nodes.hpp"
#include "gpus.hpp"
#include "env.hpp"
#include "inventory.hpp"
#include "gotypes.hpp"
#include <memory>
#include <new>
#include <stdexcept>

using namespace cro;

static_assert(sizeof(cro_probe_result) == 512, "cro_probe_result is the 512-byte all-gather payload");
static_assert(offsetof(cro_probe_result, gpu_uuid) == 16, "layout");
static_assert(offsetof(cro_probe_result, pci_bus_id) == 64, "layout");
static_assert(offsetof(cro_probe_result, checksum_xor) == 112, "layout");
static_assert(offsetof(cro_probe_result, p2p_read_ns) == 184, "layout");
static_assert(offsetof(cro_probe_result, p2p_access) == 344, "layout");
static_assert(offsetof(cro_probe_result, p2p_bytes) == 352, "layout");
static_assert(offsetof(cro_probe_result, p2p_write_ns) == 424, "layout");
static_assert(offsetof(cro_probe_result, nonce) == 488, "layout");
static_assert(offsetof(cro_probe_result, t_start_ns) == 504, "layout");
static_assert(sizeof(cro_fullbox_time) == 64 && offsetof(cro_fullbox_time, gather) == 56, "cro_fullbox_time layout");
static_assert(sizeof(cro_sweep_result) == 56, "layout");

using namespace cro::capi;


namespace cro {
namespace capi {
int on_exception() noexcept {
    try {
        throw;
    } catch (const std::bad_alloc&) {
        set_thread_error("out of host memory");
        return CRO_ERR_OOM;
    } catch (const std::exception& e) {
        set_thread_error(std::string("internal error: ") + e.what());
        return CRO_ERR_INTERNAL;
    } catch (...) {
        set_thread_error("internal error: unknown exception");
        return CRO_ERR_INTERNAL;
    }
}
}  // namespace capi
}  // namespace cro

extern "C" {

const char* cro_version(void) { return "croprobe 0.2.0 (sm_100a, abi 2)"; }

const char* cro_strerror(int code) {
    switch (code) {
        case CRO_OK: return "ok";
        case CRO_ERR_INVALID_ARG: return "invalid argument";
        case CRO_ERR_ABI_MISMATCH: return "abi version mismatch";
        case CRO_ERR_NO_DEVICE: return "no CUDA device";
        case CRO_ERR_CUDA: return "cuda runtime error";
        case CRO_ERR_OOM: return "sweep buffer allocation failed";
        case CRO_ERR_CHECKSUM: return "hbm checksum mismatch";
        case CRO_ERR_BUFFER_SMALL: return "output buffer too small";
        case CRO_ERR_NCCL: return "nccl error";
        case CRO_ERR_DEADLINE: return "deadline exceeded";
        case CRO_ERR_UNSUPPORTED: return "unsupported field";
        case CRO_ERR_PARSE: return "parse error";
        case CRO_ERR_EXEC: return "enumerate command failed";
        case CRO_ERR_P2P: return "peer access error";
        case CRO_ERR_INTERNAL: return "internal error";
        default: return "unknown error";
    }
}

int cro_last_error(cro_ctx* ctx, char* buf, size_t cap) try {
    if (!ctx) return copy_out(last_init_error(), buf, cap, nullptr);
    std::lock_guard<std::mutex> g(ctx->err_mu);
    return copy_out(ctx->last_error, buf, cap, nullptr);
} CRO_API_CATCH

int cro_selftest_exception_barrier(int kind) try {
    if (kind == 0) throw std::runtime_error("exception barrier self-test");
    if (kind == 1) throw std::bad_alloc();
    if (kind == 2) throw 42;
    return CRO_OK;
} CRO_API_CATCH

int cro_probe_init(const cro_opts* opts, cro_ctx** out) try { return ctx_create(opts, out); } CRO_API_CATCH
void cro_probe_destroy(cro_ctx* ctx) { ctx_destroy(ctx); }

int cro_device_count(cro_ctx* ctx, int* n) try {
    if (!ctx || !n) return CRO_ERR_INVALID_ARG;
    *n = (int)ctx->devs.size();
    return CRO_OK;
} CRO_API_CATCH

int cro_enumerate(cro_ctx* ctx, cro_dev_info* out, int cap, int* n) try {
    if (!ctx || !n) return CRO_ERR_INVALID_ARG;
    std::vector<cro_dev_info> inv;
    int rc = ctx_inventory(ctx, &inv);      // fresh every call: the reference re-execs nvidia-smi per reconcile
    if (rc) return rc;
    *n = (int)inv.size();
    if (*n == 0) return CRO_OK;
    if (!out || cap < *n) return CRO_ERR_BUFFER_SMALL;
    for (int i = 0; i < *n; ++i) out[i] = inv[(size_t)i];
    return CRO_OK;
} CRO_API_CATCH

int cro_node_inventory(const char* proc_root, const cro_dev_info* in_process, int n_in, cro_dev_info* out, int cap, int* n) try {
    if (!n || n_in < 0 || (n_in > 0 && !in_process)) return CRO_ERR_INVALID_ARG;
    const std::string root = proc_root && *proc_root ? proc_root : "/proc";
    std::vector<cro_dev_info> mine(in_process, in_process + n_in);
    const bool have = inventory::ProcRegistryExists(root);
    const std::vector<cro_dev_info> inv =
        inventory::Merge(mine, have, have ? inventory::FromProc(identity::ScanProc(root)) : std::vector<inventory::Seen>());
    *n = (int)inv.size();
    if (*n == 0) return CRO_OK;
    if (!out || cap < *n) return CRO_ERR_BUFFER_SMALL;
    for (int i = 0; i < *n; ++i) out[i] = inv[(size_t)i];
    return CRO_OK;
} CRO_API_CATCH

int cro_probe_uuid(cro_ctx* ctx, const char* gpu_uuid, cro_probe_result* out) try {
    return ctx_probe_uuid(ctx, gpu_uuid, out);
} CRO_API_CATCH

int cro_emit_csv(const cro_dev_info* devs, int n, const char* query, char* buf, size_t cap, size_t* len) try {
    if (!query || (n > 0 && !devs)) return CRO_ERR_INVALID_ARG;
    std::string out, err;
    int rc = identity::EmitCsv(devs, n, query, &out, &err);
    if (rc != CRO_OK) {
        copy_out(err, buf, cap, len);
        return rc;
    }
    return copy_out(out, buf, cap, len);
} CRO_API_CATCH

static int finish_parse(const identity::GpuInfoResult& r, char* buf, size_t cap, size_t* len) {
    if (r.code != CRO_OK) {
        int rc = copy_out(r.error, buf, cap, len);
        return rc == CRO_OK ? r.code : rc;
    }
    return copy_out(identity::GpuInfosToJson(r), buf, cap, len);
}

int cro_parse_gpu_csv(const char* std_out, const char* std_err, const char* exec_err, const char* query,
                      char* buf, size_t cap, size_t* len) try {
    if (!query) return CRO_ERR_INVALID_ARG;
    return finish_parse(identity::getGPUInfoFromNvidiaSmiOutput(S(std_out), S(std_err), exec_err, query),
                        buf, cap, len);
} CRO_API_CATCH

int cro_parse_proc_csv(const char* std_out, const char* std_err, const char* exec_err, const char* query,
                       char* buf, size_t cap, size_t* len) try {
    if (!query) return CRO_ERR_INVALID_ARG;
    return finish_parse(identity::getGPUInfoFromProcOutput(S(std_out), S(std_err), exec_err, query), buf,
                        cap, len);
} CRO_API_CATCH

int cro_proc_information_to_line(const char* text, char* buf, size_t cap, size_t* len) try {
    if (!text) return CRO_ERR_INVALID_ARG;
    return copy_out(identity::ProcInformationToLine(text), buf, cap, len);
} CRO_API_CATCH

int cro_check_gpu_visible(const cro_dev_info* devs, int n, const char* device_id, int* visible) try {
    if (!visible || !device_id || (n > 0 && !devs)) return CRO_ERR_INVALID_ARG;
    *visible = identity::CheckGPUVisible(devs, n, device_id) ? 1 : 0;
    return CRO_OK;
} CRO_API_CATCH

int cro_normalize(int kind, const char* in, char* buf, size_t cap, size_t* len) try {
    if (!in) return CRO_ERR_INVALID_ARG;
    std::string out;
    int rc = identity::Normalize(kind, in, &out);
    if (rc) return rc;
    return copy_out(out, buf, cap, len);
} CRO_API_CATCH

int cro_probe_device(cro_ctx* ctx, int dev_index, cro_probe_result* out) try {
    if (!ctx) return CRO_ERR_INVALID_ARG;
    return ctx_probe_device(ctx, dev_index, out);
} CRO_API_CATCH

int cro_probe_begin(cro_ctx* ctx, int dev_index) try { return ctx ? ctx_probe_begin(ctx, dev_index) : CRO_ERR_INVALID_ARG; } CRO_API_CATCH
int cro_probe_end(cro_ctx* ctx, int dev_index, cro_probe_result* out) try {
    return ctx ? ctx_probe_end(ctx, dev_index, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH

int cro_probe_all(cro_ctx* ctx, cro_probe_result* out, int cap, int* n) try {
    return ctx_probe_all(ctx, out, cap, n);
} CRO_API_CATCH

int cro_result_device_ptr(cro_ctx* ctx, int dev_index, uint64_t* dptr) try {
    if (!ctx || !dptr || dev_index < 0 || dev_index >= (int)ctx->devs.size()) return CRO_ERR_INVALID_ARG;
    *dptr = (uint64_t)(uintptr_t)ctx->devs[(size_t)dev_index]->d_result;
    return CRO_OK;
} CRO_API_CATCH

int cro_hbm_fill(cro_ctx* ctx, int i, cro_sweep_result* out) try { return ctx ? ctx_fill(ctx, i, 1, out) : CRO_ERR_INVALID_ARG; } CRO_API_CATCH
int cro_hbm_fill_loop(cro_ctx* ctx, int i, uint32_t iters, cro_sweep_result* out) try {
    return ctx ? ctx_fill(ctx, i, iters, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_hbm_read_checksum(cro_ctx* ctx, int i, uint32_t variant, cro_sweep_result* out) try {
    return ctx ? ctx_read(ctx, i, variant, 1, false, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_hbm_read_checksum_dst(cro_ctx* ctx, int i, uint32_t variant, cro_sweep_result* out) try {
    return ctx ? ctx_read(ctx, i, variant, 1, true, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_hbm_read_loop(cro_ctx* ctx, int i, uint32_t variant, uint32_t iters, cro_sweep_result* out) try {
    return ctx ? ctx_read(ctx, i, variant, iters, false, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_hbm_copy(cro_ctx* ctx, int i, uint32_t variant, cro_sweep_result* out) try {
    return ctx ? ctx_copy(ctx, i, variant, 1, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_hbm_copy_loop(cro_ctx* ctx, int i, uint32_t variant, uint32_t iters, cro_sweep_result* out) try {
    return ctx ? ctx_copy(ctx, i, variant, iters, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_hbm_expected_checksum(cro_ctx* ctx, int i, cro_sweep_result* out) try {
    return ctx ? ctx_expected(ctx, i, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_inject_fault(cro_ctx* ctx, int i, uint64_t word, uint64_t mask) try {
    return ctx ? ctx_inject(ctx, i, word, mask) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_read_words(cro_ctx* ctx, int i, uint64_t first, uint64_t n, uint64_t* out) try {
    return ctx ? ctx_read_words(ctx, i, first, n, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_device_seed(cro_ctx* ctx, int i, uint64_t* seed) try {
    if (!ctx || !seed || i < 0 || i >= (int)ctx->devs.size()) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(ctx->devs[(size_t)i]->mu);
    *seed = ctx->devs[(size_t)i]->seed_cur;
    return CRO_OK;
} CRO_API_CATCH

int cro_probe_sweep_times(cro_ctx* ctx, int i, cro_sweep_time* out, int cap, int* n) try {
    return ctx ? ctx_sweep_times(ctx, i, out, cap, n) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH

int cro_p2p_detail_get(cro_ctx* ctx, int i, int peer, cro_p2p_detail* out) try {
    return ctx ? ctx_p2p_detail(ctx, i, peer, out) : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH

int cro_fullbox_times(cro_ctx* ctx, cro_fullbox_time* out) try {
    if (!ctx || !out) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(ctx->all_mu);
    const FullBoxTimes& f = ctx->fullbox;
    out->enqueue_ns = f.enqueue_ns;
    out->wall_ns = f.wall_ns;
    out->hbm_ns = f.hbm_ns;
    out->p2p_ns = f.p2p_ns;
    out->chase_ns = f.chase_ns;
    out->gather_ns = f.gather_ns;
    out->rounds = f.rounds;
    out->host_syncs = f.host_syncs;
    out->gather = f.gather;
    out->reserved = 0;
    return CRO_OK;
} CRO_API_CATCH

int cro_set_latency_hops(cro_ctx* ctx, uint32_t hops) try {
    if (!ctx || hops == 0 || hops > (1u << 24)) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(ctx->all_mu);
    ctx->opts.latency_hops = hops;      // the next cro_probe_all recomputes where each chase must end
    return CRO_OK;
} CRO_API_CATCH

int cro_metrics_text(cro_ctx* ctx, char* buf, size_t cap, size_t* len) try {
    if (!ctx) return CRO_ERR_INVALID_ARG;
    return copy_out(ctx_metrics_text(ctx), buf, cap, len);
} CRO_API_CATCH

static void describe_type(const gojson::GoType& t, gojson::Writer* w) {
    w->begin_object();
    w->field("type", t.name);
    if (t.kind == gojson::GoType::Struct) {
        w->field("struct", t.structName);
        w->key("fields").begin_array();
        for (const auto& f : t.fields) {
            w->begin_object();
            w->field("json", f.first);
            w->key("of");
            describe_type(*f.second, w);
            w->end_object();
        }
        w->end_array();
    } else if (t.kind == gojson::GoType::Slice && t.elem) {
        w->key("elem");
        describe_type(*t.elem, w);
    }
    w->end_object();
}

int cro_describe_wire_type(const char* name, char* buf, size_t cap, size_t* len) try {
    const std::string n = S(name);
    const gojson::GoType* t = n == "FMScaleUpResponse" ? &gotypes::FMScaleUpResponse()
                              : n == "FMGetMachineResponse" ? &gotypes::FMGetMachineResponse()
                              : n == "CMMachineData" ? &gotypes::CMMachineData() : nullptr;
    if (!t) return CRO_ERR_INVALID_ARG;
    gojson::Writer w;
    describe_type(*t, &w);
    return copy_out(w.take(), buf, cap, len);
} CRO_API_CATCH

int cro_chase_end(int minor_src, int minor_dst, uint32_t hops, uint32_t* end) try {
    if (!end) return CRO_ERR_INVALID_ARG;
    std::vector<uint32_t> perm;
    chase_permutation(minor_src, minor_dst, &perm);
    uint32_t at = 0;
    for (uint32_t h = 0; h < hops; ++h) at = perm[at];
    *end = at;
    return CRO_OK;
} CRO_API_CATCH

int cro_validate_env(const char* name, const char* value, char* err_buf, size_t err_cap) try {
    int n = 0;
    const env::Knob* t = env::table(&n);
    std::string why;
    if (name) {
        for (int i = 0; i < n; ++i) {
            if (S(name) != t[i].name) continue;
            unsigned v = 0;
            if (env::parse(t[i], value, &v, &why)) return CRO_OK;
            copy_out(why, err_buf, err_cap, nullptr);
            return CRO_ERR_INVALID_ARG;
        }
        copy_out("unknown knob " + S(name), err_buf, err_cap, nullptr);
        return CRO_ERR_UNSUPPORTED;
    }
    if (env::reload(&why)) return CRO_OK;
    copy_out(why, err_buf, err_cap, nullptr);
    return CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
uint64_t cro_launch_count(cro_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

// ---- emitters --------------------------------------------------------------

int cro_emit_status_json(const char* state, const char* error, const char* device_id,
                         const char* cdi_device_id, char* buf, size_t cap, size_t* len) try {
    controller::ComposableResourceStatus st;
    st.State = S(state);
    st.Error = S(error);
    st.DeviceID = S(device_id);
    st.CDIDeviceID = S(cdi_device_id);
    return copy_out(st.MarshalJSON(), buf, cap, len);
} CRO_API_CATCH

int cro_emit_scalar_status_json(const char* state, const char* device_id, const char* cdi_device_id,
                                const char* node_name, const char* error, char* buf, size_t cap,
                                size_t* len) try {
    // api/v1alpha1/composabilityrequest_types.go:74-80 (declaration order)
    gojson::Writer w;
    w.begin_object();
    w.field("state", S(state));
    w.field_omitempty("device_id", S(device_id));
    w.field_omitempty("cdi_device_id", S(cdi_device_id));
    w.field_omitempty("node_name", S(node_name));
    w.field_omitempty("error", S(error));
    w.end_object();
    return copy_out(w.str(), buf, cap, len);
} CRO_API_CATCH

int cro_emit_fm_scale_up(const char* tenant_uuid, const char* mach_uuid, const char* res_type,
                         const char* model, char* buf, size_t cap, size_t* len) try {
    return copy_out(fabric::FMScaleUpBody(S(tenant_uuid), S(mach_uuid), S(res_type), S(model)), buf, cap, len);
} CRO_API_CATCH

int cro_emit_fm_scale_down(const char* tenant_uuid, const char* mach_uuid, const char* res_type,
                           const char* res_uuid, char* buf, size_t cap, size_t* len) try {
    return copy_out(fabric::FMScaleDownBody(S(tenant_uuid), S(mach_uuid), S(res_type), S(res_uuid)), buf, cap, len);
} CRO_API_CATCH

int cro_emit_cm_scale_up(const char* spec_uuid, int device_count, char* buf, size_t cap, size_t* len) try {
    return copy_out(fabric::CMScaleUpBody(S(spec_uuid), device_count), buf, cap, len);
} CRO_API_CATCH

int cro_emit_cm_scale_down(const char* spec_uuid, int device_count, const char* device_id, char* buf,
                           size_t cap, size_t* len) try {
    return copy_out(fabric::CMScaleDownBody(S(spec_uuid), device_count, S(device_id)), buf, cap, len);
} CRO_API_CATCH

int cro_emit_sunfish_request(const char* name, long long count, const char* proc_type, const char* model,
                             char* buf, size_t cap, size_t* len) try {
    return copy_out(fabric::SunfishBody(S(name), count, S(proc_type), S(model)), buf, cap, len);
} CRO_API_CATCH

}  // extern "C" (reopened below)
std::map<std::string, std::string> cro::capi::probe_annotations(const cro_probe_result& r) {
    std::map<std::string, std::string> m;
    m["cohdi.io/probe-status"] = cro_strerror(r.status);
    m["cohdi.io/probe-device-id"] = fixed_str(r.gpu_uuid, sizeof r.gpu_uuid);
    m["cohdi.io/probe-pci-bus-id"] = fixed_str(r.pci_bus_id, sizeof r.pci_bus_id);
    m["cohdi.io/probe-device-minor"] = std::to_string(r.device_minor);
    m["cohdi.io/probe-sweep-bytes"] = std::to_string(r.sweep_bytes);
    m["cohdi.io/probe-checksum"] = hex16(r.checksum_xor) + ":" + hex16(r.checksum_sum) + ":" + hex16(r.checksum_wsum);
    m["cohdi.io/probe-nonce"] = std::to_string(r.nonce);
    if (r.copy_sweeps) m["cohdi.io/probe-copies-verified"] = std::to_string((unsigned)r.copy_verified) + "/" + std::to_string((unsigned)r.copy_sweeps);
    m["cohdi.io/probe-ecc-uncorrected"] = std::to_string(r.ecc_errors);
    m["cohdi.io/probe-hbm-fill-gbs"] = gbs_x10(r.sweep_bytes, r.fill_ns);
    m["cohdi.io/probe-hbm-read-gbs"] = gbs_x10(r.sweep_bytes, r.read_best_ns);
    if (r.copy_sweeps) m["cohdi.io/probe-hbm-copy-gbs"] = gbs_x10(2 * r.sweep_bytes, r.copy_best_ns);
    std::string bw, lat;
    for (int j = 0; j < 8; ++j) {
        if (!r.p2p_read_ns[j]) continue;
        if (!bw.empty()) { bw += ","; lat += ","; }
        bw += std::to_string(j) + ":" + gbs_x10(r.p2p_bytes, r.p2p_read_ns[j]);
        lat += std::to_string(j) + ":" + std::to_string(r.p2p_latency_ns_x16[j] / 16);
    }
    if (!bw.empty()) {
        m["cohdi.io/probe-nvlink-read-gbs"] = bw;
        m["cohdi.io/probe-nvlink-latency-ns"] = lat;
    }
    std::string wr;
    for (int j = 0; j < 8; ++j) {
        if (!r.p2p_write_ns[j]) continue;
        if (!wr.empty()) wr += ",";
        wr += std::to_string(j) + ":" + gbs_x10(r.p2p_bytes, r.p2p_write_ns[j]);
    }
    if (!wr.empty()) m["cohdi.io/probe-nvlink-write-gbs"] = wr;
    return m;
}

extern "C" {

int cro_emit_probe_annotations_json(const cro_probe_result* r, char* buf, size_t cap, size_t* len) try {
    if (!r) return CRO_ERR_INVALID_ARG;
    gojson::Writer w;
    w.string_map(probe_annotations(*r));
    return copy_out(w.str(), buf, cap, len);
} CRO_API_CATCH

int cro_fm_parse_scale_up_response(const char* body, const char* resource_name, const char* res_type,
                                   const char* model, char* device_id, size_t device_id_cap,
                                   char* cdi_device_id, size_t cdi_cap, char* err_buf, size_t err_cap) try {
    if (!body) return CRO_ERR_INVALID_ARG;
    std::string dev, cdi;
    controller::Error e = controller::FMScaleUpResponseToIDs(body, S(resource_name), S(res_type), S(model), &dev, &cdi);
    if (!e.ok()) {
        copy_out(e.msg, err_buf, err_cap, nullptr);
        return CRO_ERR_PARSE;
    }
    int rc = copy_out(dev, device_id, device_id_cap, nullptr);
    if (rc) return rc;
    return copy_out(cdi, cdi_device_id, cdi_cap, nullptr);
} CRO_API_CATCH

int cro_cm_check_adding_resources(const char* machine_body, const char* existing_device_ids,
                                  const char* res_type, const char* model, char* spec_uuid, size_t spec_cap,
                                  int* device_count, char* device_id, size_t device_id_cap,
                                  char* cdi_device_id, size_t cdi_cap, char* err_buf, size_t err_cap) try {
    if (!machine_body) return CRO_ERR_INVALID_ARG;
    std::vector<std::string> existing;
    if (existing_device_ids) existing = identity::Split(existing_device_ids, "\n");
    controller::CMAddingResult r = controller::CMCheckAddingResources(machine_body, existing, S(res_type), S(model));
    if (device_count) *device_count = (int)r.deviceCount;
    int rc = copy_out(r.specUUID, spec_uuid, spec_cap, nullptr);
    if (rc) return rc;
    if ((rc = copy_out(r.deviceID, device_id, device_id_cap, nullptr))) return rc;
    if ((rc = copy_out(r.CDIDeviceID, cdi_device_id, cdi_cap, nullptr))) return rc;
    copy_out(r.err.ok() ? std::string() : r.err.msg, err_buf, err_cap, nullptr);
    return r.err.ok() ? CRO_OK : CRO_ERR_PARSE;
} CRO_API_CATCH

// ---- fabric wire codec ---------------------------------------------------------------

int cro_fabric_check_resource(const char* kind, const char* machine_body, const char* res_type, const char* model,
                              const char* device_id, char* err_buf, size_t err_cap) try {
    if (!kind || !machine_body) return CRO_ERR_INVALID_ARG;
    controller::Error e;
    if (S(kind) == "fm") e = fabric::FMCheckResource(machine_body, S(res_type), S(model), S(device_id));
    else if (S(kind) == "cm") e = fabric::CMCheckResource(machine_body, S(res_type), S(model), S(device_id));
    else return CRO_ERR_INVALID_ARG;
    copy_out(e.ok() ? std::string() : e.msg, err_buf, err_cap, nullptr);
    return e.ok() ? CRO_OK : CRO_ERR_EXEC;
} CRO_API_CATCH

int cro_fabric_get_resources(const char* kind, const char* machine_body, const char* node_name, const char* machine_uuid,
                             char* buf, size_t cap, size_t* len) try {
    if (!kind || !machine_body) return CRO_ERR_INVALID_ARG;
    std::vector<fabric::DeviceInfo> v;
    controller::Error e;
    if (S(kind) == "fm") e = fabric::FMGetResources(machine_body, S(node_name), S(machine_uuid), &v);
    else if (S(kind) == "cm") e = fabric::CMGetResources(machine_body, S(node_name), S(machine_uuid), &v);
    else return CRO_ERR_INVALID_ARG;
    if (!e.ok()) {
        copy_out(e.msg, buf, cap, len);
        return CRO_ERR_PARSE;
    }
    return copy_out(fabric::DeviceInfosToJson(v), buf, cap, len);
} CRO_API_CATCH

// ---- detach-side pre-flight -------------------------------------------------------

static int finish_err(const controller::Error& e, char* err_buf, size_t err_cap) {
    copy_out(e.ok() ? std::string() : e.msg, err_buf, err_cap, nullptr);
    return e.ok() ? CRO_OK : CRO_ERR_EXEC;
}

int cro_check_no_gpu_loads(const char* std_out, const char* std_err, const char* exec_err, const char* pod_name,
                           const char* node_name, const char* target_uuid, int driver_enabled, char* err_buf,
                           size_t err_cap) try {
    std::string uuid = S(target_uuid);
    return finish_err(detach::CheckNoGPULoadsFromOutput(S(std_out), S(std_err), exec_err, S(pod_name), S(node_name),
                                                        target_uuid ? &uuid : nullptr, driver_enabled != 0),
                      err_buf, err_cap);
} CRO_API_CATCH

int cro_check_gpu_drain_status(const char* std_out, const char* std_err, const char* exec_err, const char* node_name,
                               const char* bus_id, int* draining, char* err_buf, size_t err_cap) try {
    bool d = false;
    controller::Error e = detach::checkGPUDrainStatusFromOutput(S(std_out), S(std_err), exec_err, S(node_name), S(bus_id), &d);
    if (draining) *draining = d ? 1 : 0;
    return finish_err(e, err_buf, err_cap);
} CRO_API_CATCH

int cro_check_device_file_scan(const char* std_out, const char* std_err, const char* exec_err, int rke2, char* err_buf,
                               size_t err_cap) try {
    return finish_err(detach::CheckDeviceFileScanResult(S(std_out), S(std_err), exec_err, rke2 != 0), err_buf, err_cap);
} CRO_API_CATCH

int cro_scan_device_file_holders(const char* proc_root, const char* target, int rke2, char* buf, size_t cap, size_t* len) try {
    if (!target) return CRO_ERR_INVALID_ARG;
    return copy_out(detach::ScanDeviceFileHolders(S(proc_root), target, rke2 != 0), buf, cap, len);
} CRO_API_CATCH

// ---- in-memory cluster -----------------------------------------------------------

struct cro_sim {
    std::unique_ptr<sim::Cluster> cluster;
    std::mutex mu;
};

int cro_sim_create(cro_ctx* ctx, const char* config_json, cro_sim** out) try {
    if (!out) return CRO_ERR_INVALID_ARG;
    std::string perr;
    gojson::ValuePtr cfg = gojson::parse(config_json ? config_json : "{}", &perr);
    if (!cfg || cfg->kind != gojson::Value::Object) return CRO_ERR_PARSE;
    cro_sim* s = new cro_sim;
    s->cluster.reset(new sim::Cluster(ctx, *cfg));
    *out = s;
    return CRO_OK;
} CRO_API_CATCH
void cro_sim_destroy(cro_sim* s) { delete s; }

static int sim_json_call(cro_sim* s, const char* json, char* err_buf, size_t err_cap,
                         controller::Error (sim::Cluster::*fn)(const gojson::Value&)) {
    if (!s || !json) return CRO_ERR_INVALID_ARG;
    std::string perr;
    gojson::ValuePtr v = gojson::parse(json, &perr);
    if (!v || v->kind != gojson::Value::Object) {
        copy_out(perr, err_buf, err_cap, nullptr);
        return CRO_ERR_PARSE;
    }
    std::lock_guard<std::mutex> g(s->mu);
    controller::Error e = (s->cluster.get()->*fn)(*v);
    copy_out(e.ok() ? std::string() : e.msg, err_buf, err_cap, nullptr);
    return e.ok() ? CRO_OK : CRO_ERR_INVALID_ARG;
}
int cro_sim_apply(cro_sim* s, const char* json, char* err_buf, size_t err_cap) try {
    return sim_json_call(s, json, err_buf, err_cap, &sim::Cluster::Apply);
} CRO_API_CATCH
int cro_sim_plant(cro_sim* s, const char* json, char* err_buf, size_t err_cap) try {
    return sim_json_call(s, json, err_buf, err_cap, &sim::Cluster::Plant);
} CRO_API_CATCH
int cro_sim_delete(cro_sim* s, const char* name) try {
    if (!s || !name) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(s->mu);
    return s->cluster->Delete(name).ok() ? CRO_OK : CRO_ERR_INVALID_ARG;
} CRO_API_CATCH
int cro_sim_run(cro_sim* s, long long max_reconciles, char* buf, size_t cap, size_t* len) try {
    if (!s) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(s->mu);
    s->cluster->Run(max_reconciles > 0 ? max_reconciles : (1ll << 40));
    return copy_out(s->cluster->StatsJSON(), buf, cap, len);
} CRO_API_CATCH
int cro_sim_reconcile_request(cro_sim* s, const char* name, char* err_buf, size_t err_cap) try {
    if (!s || !name) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(s->mu);
    controller::Error e = s->cluster->ReconcileRequestOnce(name);
    copy_out(e.ok() ? std::string() : e.msg, err_buf, err_cap, nullptr);
    return e.ok() ? CRO_OK : CRO_ERR_EXEC;
} CRO_API_CATCH
int cro_sim_reconcile_resource(cro_sim* s, const char* name, char* err_buf, size_t err_cap) try {
    if (!s || !name) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(s->mu);
    controller::Error e = s->cluster->ReconcileResourceOnce(name);
    copy_out(e.ok() ? std::string() : e.msg, err_buf, err_cap, nullptr);
    return e.ok() ? CRO_OK : CRO_ERR_EXEC;
} CRO_API_CATCH
int cro_sim_sync_upstream(cro_sim* s, const char* devices_json, long long now_s, char* err_buf, size_t err_cap) try {
    if (!s || !devices_json) return CRO_ERR_INVALID_ARG;
    std::string perr;
    gojson::ValuePtr v = gojson::parse(devices_json, &perr);
    if (!v) {
        copy_out("failed to fetch data from upstream server: " + perr, err_buf, err_cap, nullptr);
        return CRO_ERR_PARSE;
    }
    std::lock_guard<std::mutex> g(s->mu);
    controller::Error e = s->cluster->SyncUpstream(*v, now_s);
    copy_out(e.ok() ? std::string() : e.msg, err_buf, err_cap, nullptr);
    return e.ok() ? CRO_OK : CRO_ERR_EXEC;
} CRO_API_CATCH
int cro_sim_dump(cro_sim* s, char* buf, size_t cap, size_t* len) try {
    if (!s) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(s->mu);
    return copy_out(s->cluster->DumpJSON(), buf, cap, len);
} CRO_API_CATCH

}  // extern "C"
