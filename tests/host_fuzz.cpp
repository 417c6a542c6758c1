// Deterministic mutation fuzzer for the host-side parsers of libcroprobe,
// built with -fsanitize=address,undefined by tests/test_host_sanitizers.py.
// Everything these functions read comes from outside the operator (exec
// output of nvidia-smi / awk, HTTP bodies of the fabric managers), so none of
// them may read out of bounds, overflow or leak on arbitrary bytes.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "detach.hpp"
#include "fabric.hpp"
#include "gojson.hpp"
#include "gpus.hpp"
#include "identity.hpp"
#include "nodes.hpp"
#include "nvml_ops.hpp"
#include "provider.hpp"
#include "reconcile.hpp"

using namespace cro;

static thread_local uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

static const char* kCorpus[] = {
    // nvidia-smi csv,noheader
    "0, GPU-7cc45b7b-2a6d-f0ac-1b02-6f8de09e1a6c, 00000000:1F:00.0\n1, GPU-aaaa, 00000000:20:00.0\n",
    "No devices were found\n",
    "GPU-1, python3\nGPU-2, pytorch\n",
    "",
    "0000:1F:00.0", "00000000:40:00.0", "40:00",
    "\n\n , ,\n",
    // awk over /proc/driver/nvidia/gpus/*/information
    "Model: NVIDIA B200\nGPU UUID: GPU-7cc45b7b\nDevice Minor: 3\nBus Location: 0000:1f:00.0\n",
    "3,GPU-7cc45b7b,0000:1f:00.0\n",
    // drain status
    "GPU 0000:1F:00.0 is currently: draining\n",
    "GPU 0000:1F:00.0 is currently: not draining\n",
    // FM scale-up response / machine
    "{\"data\":{\"machines\":[{\"fabric_uuid\":\"f\",\"fabric_id\":1,\"mach_uuid\":\"m\",\"mach_id\":1,\"mach_name\":\"n\","
    "\"tenant_uuid\":\"t\",\"resources\":[{\"res_uuid\":\"GPU-1\",\"res_name\":\"res-0\",\"res_type\":\"gpu\","
    "\"res_status\":1,\"res_op_status\":\"0\",\"res_serial_num\":\"s\",\"res_spec\":{\"condition\":[{\"column\":\"model\","
    "\"operator\":\"eq\",\"value\":\"NVIDIA-B200\"}]}}]}]}}",
    // CM machine
    "{\"data\":{\"cluster\":{\"cluster_uuid\":\"c\",\"machine\":{\"uuid\":\"m\",\"name\":\"n\",\"status\":\"s\","
    "\"status_reason\":\"\",\"resspecs\":[{\"spec_uuid\":\"su\",\"type\":\"gpu\",\"selector\":{\"version\":\"1\","
    "\"expression\":{\"conditions\":[{\"column\":\"model\",\"operator\":\"eq\",\"value\":\"NVIDIA-B200\"}]}},"
    "\"min_resspec_count\":0,\"max_resspec_count\":2,\"device_count\":1,\"devices\":[{\"device_uuid\":\"GPU-1\","
    "\"status\":\"ADD_COMPLETE\",\"status_reason\":\"\",\"detail\":{\"fabr_gid\":\"g\",\"res_uuid\":\"GPU-1\","
    "\"fabr_uuid\":\"f\",\"res_type\":\"gpu\",\"res_name\":\"r\",\"res_status\":\"1\",\"res_op_status\":\"0\","
    "\"res_spec\":[{\"name\":\"model\",\"value\":\"NVIDIA-B200\"}],\"tenant_id\":\"t\",\"mach_id\":\"m\"}}]}]}}}}",
    "{\"a\":[1,2.5e3,-0,true,false,null,\"\\u00e9\\ud83d\\ude00\\n\"],\"b\":{\"c\":{}}}",
    "[[[[[[[[[[[[[[[[[[[[[[[[[[[[[[[[]]]]]]]]]]]]]]]]]]]]]]]]]]]]]]]]",
    // fabric error bodies and timestamps
    "{\"status\":404,\"detail\":{\"code\":\"E02XXXX\",\"message\":\"machine not found\"}}",
    "{\"status\":404.5,\"detail\":{\"code\":7,\"message\":{\"k\": [1, 2]},\"data\":[]}}",
    "<html><body>This is not JSON!</body></html>",
    "2025-06-01T12:00:00Z", "2024-02-29T23:59:59.123456789+09:00", "0000-01-01T00:00:00-24:00", "9999-12-31T23:59:59,5Z",
};

static std::string mutate(std::string s) {
    const int ops = 1 + (int)(rnd() % 4);
    for (int k = 0; k < ops; ++k) {
        const uint64_t r = rnd();
        switch (r % 7) {
        case 0: if (!s.empty()) s[(r >> 8) % s.size()] = (char)(r >> 40); break;                   // byte flip
        case 1: if (!s.empty()) s.resize((r >> 8) % s.size()); break;                               // truncate
        case 2: s.insert((r >> 8) % (s.size() + 1), 1, (char)(r >> 40)); break;                     // insert
        case 3: if (!s.empty()) { size_t a = (r >> 8) % s.size(); s += s.substr(a, (r >> 32) % 64); } break;  // splice
        case 4: if (!s.empty()) s.erase((r >> 8) % s.size(), (r >> 32) % 8); break;                 // delete
        case 5: { static const char* tok[] = {"\"", "\\", "\\u", "\\ud800", "{", "[", ",", ":", "\n", "\r\n", ", ",
                                               "\xc3", "\xe2\x80\xa8", "\xff", "1e999", "-", "0000", "null"};
                  s.insert((r >> 8) % (s.size() + 1), tok[(r >> 40) % (sizeof tok / sizeof *tok)]); } break;
        default: { const char* o = kCorpus[(r >> 8) % (sizeof kCorpus / sizeof *kCorpus)]; s += o; } break;
        }
    }
    return s;
}

static thread_local size_t sink = 0;   // keeps results alive

// Node-side flows (csrc/gpus.cpp) with a pod-exec that answers every request with the fuzz input:
// whatever nvidia-smi / lsmod / the scans print, the decision code must stay in bounds.
struct FuzzKube : gpus::Kube {
    int policy = 0;   // 0 NotFound, 1 unset, 2 disabled, 3 enabled
    controller::Error GetClusterPolicy(bool* found, bool* set, bool* enabled) override {
        *found = policy > 0; *set = policy > 1; *enabled = policy == 3;
        return controller::Error::Nil();
    }
    controller::Error ListPods(std::vector<gpus::Pod>* out) override {
        out->push_back({"ns", "nvidia-driver-daemonset-x", "worker-0", {{"app.kubernetes.io/component", "nvidia-driver"}}, {"ctr"}});
        out->push_back({"ns", "nvidia-dra-driver-gpu-kubelet-plugin-x", "worker-0", {{"app.kubernetes.io/name", "nvidia-dra-driver-gpu"}}, {}});
        out->push_back({"ns", "cro-node-agent-x", "worker-0", {{"app", "cro-node-agent"}}, {"agent"}});
        return controller::Error::Nil();
    }
    controller::Error ListResourceSliceUUIDs(std::vector<std::string>* out) override { out->push_back("GPU-1"); return controller::Error::Nil(); }
};
struct FuzzExec : gpus::Exec {
    const std::string* text = nullptr;
    unsigned n = 0;
    gpus::ExecResult Run(const gpus::Pod&, const std::string& container, const gpus::ExecRequest& req) override {
        gpus::ExecResult r;
        ++n;
        sink += gpus::ExecRawQuery(req.kind == gpus::ExecRequest::Command ? req.argv : gpus::ScanAsCommand(req), container).size();
        if (n % 3 != 0) r.std_out = *text;          // sometimes the answer is on stdout, sometimes stderr, sometimes empty
        else if (n % 2 == 0) r.std_err = *text;
        if (n % 11 == 0) { r.failed = true; r.exec_err = "command terminated with exit code 1"; }
        return r;
    }
};

static void one(const std::string& in, const std::string& err_text) {
    const char* exec_err = (rnd() & 7) == 0 ? "command terminated with exit code 1" : nullptr;
    static const char* queries[] = {"device_minor,gpu_uuid,pci.bus_id", "gpu_uuid", "index,gpu_uuid,name,pci.bus_id", ""};
    const std::string q = queries[rnd() % 4];
    {
        auto r = identity::getGPUInfoFromNvidiaSmiOutput(in, err_text, exec_err, q);
        sink += identity::GpuInfosToJson(r).size() + r.error.size();
        auto p = identity::getGPUInfoFromProcOutput(in, err_text, exec_err, q);
        sink += identity::GpuInfosToJson(p).size() + p.error.size();
        sink += identity::ProcInformationToLine(in).size();
        std::string out;
        for (int kind = 0; kind < 6; ++kind) sink += (size_t)identity::Normalize(kind, in, &out) + out.size();
        sink += identity::TrimSpace(in).size() + identity::Split(in, q.empty() ? "," : ", ").size();
        unsigned domain = 0, bus = 0, device = 0;          // the `-p <bus>` argument of the drain commands
        if (nvml::ParseBusId(in, &domain, &bus, &device)) sink += domain + bus + device;
    }
    {
        const std::string target = "GPU-1";
        sink += detach::CheckNoGPULoadsFromOutput(in, err_text, exec_err, "pod", "node", (rnd() & 1) ? &target : nullptr,
                                                  rnd() & 1).msg.size();
        bool draining = false;
        sink += detach::checkGPUDrainStatusFromOutput(in, err_text, exec_err, "node", "0000:1F:00.0", &draining).msg.size();
        sink += detach::CheckDeviceFileScanResult(in, err_text, exec_err, rnd() & 1).msg.size();
    }
    {
        std::string e;
        auto v = gojson::parse(in, &e);
        sink += e.size() + (v ? v->arr.size() + v->obj.size() : 0);
        std::string quoted;
        gojson::append_string(quoted, in);
        std::string e2;
        auto back = gojson::parse(quoted, &e2);          // whatever goes in must come back as a string value
        if (!back || back->kind != gojson::Value::String) {
            std::fprintf(stderr, "append_string produced unparsable JSON: %s\n", e2.c_str());
            std::abort();
        }
    }
    {
        std::string id, cdi;
        sink += controller::FMScaleUpResponseToIDs(in, "res-0", "gpu", "NVIDIA-B200", &id, &cdi).msg.size() + id.size();
        auto a = controller::CMCheckAddingResources(in, {"GPU-0"}, "gpu", "NVIDIA-B200");
        sink += sizeof a;
        std::vector<fabric::DeviceInfo> devs;
        sink += fabric::FMCheckResource(in, "gpu", "NVIDIA-B200", "GPU-1").msg.size();
        sink += fabric::CMCheckResource(in, "gpu", "NVIDIA-B200", "GPU-1").msg.size();
        sink += fabric::FMGetResources(in, "node", "m", &devs).msg.size();
        sink += fabric::CMGetResources(in, "node", "m", &devs).msg.size();
        sink += fabric::DeviceInfosToJson(devs).size();
        // non-200 replies and the detach-side machine scan
        sink += fabric::FMErrorFromReply("scaleup", in).msg.size() + fabric::FMErrorFromReply("scaledown", in).msg.size();
        sink += fabric::CMErrorFromReply("get", in).msg.size() + fabric::CMErrorFromReply("scaledown", in).msg.size();
        sink += fabric::CMCheckRemovingResources(in, "gpu", "NVIDIA-B200", "GPU-1").specUUID.size();
    }
    {
        // the id_manager's answer: as the reply body, and as the access token's middle part
        long long exp = 0;
        fabric::TokenReply tr;
        tr.body = in;
        sink += fabric::TokenFromReply(tr, &exp).msg.size();
        std::string dec, derr;
        sink += (size_t)fabric::DecodeBase64RawURL(in, &dec, &derr) + dec.size() + derr.size();
        std::string quoted;
        gojson::append_string(quoted, in);
        tr.body = "{\"access_token\":" + quoted + "}";
        sink += fabric::TokenFromReply(tr, &exp).msg.size();
        static const char kAlpha[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_";
        std::string mid;                       // RawURL-encode `in` so the claims decoder sees the mutated text too
        for (size_t i = 0; i < in.size(); i += 3) {
            const unsigned a = (unsigned char)in[i], b = i + 1 < in.size() ? (unsigned char)in[i + 1] : 0,
                           c = i + 2 < in.size() ? (unsigned char)in[i + 2] : 0;
            mid.push_back(kAlpha[a >> 2]);
            mid.push_back(kAlpha[((a & 3) << 4) | (b >> 4)]);
            if (i + 1 < in.size()) mid.push_back(kAlpha[((b & 15) << 2) | (c >> 6)]);
            if (i + 2 < in.size()) mid.push_back(kAlpha[c & 63]);
        }
        if (!fabric::DecodeBase64RawURL(mid, &dec, &derr) || dec != in) {
            std::fprintf(stderr, "base64 round trip failed\n");
            std::abort();
        }
        tr.body = "{\"access_token\":\"h." + mid + ".s\"}";
        fabric::ReplyTokenSource src(tr, 1748779200);
        sink += src.GetToken().msg.size() + src.GetToken().msg.size() + (size_t)src.fetches;
    }
    {
        // Go's validity scanner must agree with the tree builder on what is valid JSON
        const std::string syn = gojson::SyntaxError(in);
        std::string e;
        const bool parsed = (bool)gojson::parse(in, &e);
        if (syn.empty() != parsed) {
            std::fprintf(stderr, "scanner and parser disagree: syntax='%s' parse='%s'\n", syn.c_str(), e.c_str());
            std::abort();
        }
        long long t = 0, ns = 0;
        std::string terr;
        if (nodes::ParseRFC3339(in, &t, &ns, &terr)) {
            const std::string back = nodes::FormatRFC3339UTC(t);      // a parsed instant formats and re-parses to itself
            long long t2 = 0, ns2 = 0;
            if (!nodes::ParseRFC3339(back, &t2, &ns2, &terr) || t2 != t) {
                std::fprintf(stderr, "RFC3339 round trip failed: %s -> %s\n", in.c_str(), back.c_str());
                std::abort();
            }
        }
        nodes::DaemonSetView ds;
        ds.DesiredNumberScheduled = ds.NumberReady = ds.CurrentNumberScheduled = 1;
        ds.hasRestartedAt = true;
        ds.restartedAt = in;
        nodes::Restart what;
        sink += nodes::RestartDaemonsetDecision("ns", "ds", ds, 1750000000, 0, &what).msg.size();
    }
    {
        FuzzKube kube;
        FuzzExec exec;
        exec.text = &in;
        gpus::GpuNodeOps ops(&kube, &exec);
        controller::ComposableResource res;
        res.Spec.TargetNode = "worker-0";
        res.Status.DeviceID = "GPU-1";
        const std::string uuid = "GPU-1";
        for (kube.policy = 0; kube.policy < 4; ++kube.policy)
            for (const char* type : {"DRA", "DEVICE_PLUGIN"}) {
                bool visible = false;
                sink += ops.RunNvidiaSmi("worker-0").msg.size() + ops.CheckGPUVisible(type, res, &visible).msg.size();
                sink += ops.CheckNoGPULoadsFor("worker-0", &uuid).msg.size() + ops.CheckNoGPULoadsFor("worker-0", nullptr).msg.size();
                sink += ops.DrainGPU("worker-0", uuid, type).msg.size();
            }
    }
}

int main(int argc, char** argv) {
    const long iters = argc > 1 ? std::atol(argv[1]) : 20000;
    const size_t n = sizeof kCorpus / sizeof *kCorpus;
    for (size_t i = 0; i < n; ++i) one(kCorpus[i], "");
    one(std::string(1 << 20, '['), "");                      // nesting bomb: must fail cleanly, not overflow the stack
    {
        // valid documents nested deeper than the tree builder recurses (but within Go's 10000): they must parse, as an
        // unknown field, as a RawMessage and as map[string]any content alike
        const std::string deep = std::string(9000, '[') + "{\"k\":\"]}\\\"[\"}" + std::string(9000, ']');
        one(deep, "");
        one("{\"status\":500,\"detail\":{\"code\":\"E\",\"message\":" + deep + ",\"data\":{\"x\":" + deep + "}},\"junk\":" + deep + "}", "");
        one("{\"data\":{\"machines\":[{\"unknown\":" + deep + ",\"resources\":[]}]}}", "");
        std::string e;
        if (!gojson::parse(deep, &e)) {
            std::fprintf(stderr, "deep but valid JSON was refused: %s\n", e.c_str());
            std::abort();
        }
    }
    one("{\"a\":" + std::string(1 << 18, '{'), "");
    one(std::string(1 << 20, ','), "");                      // a million empty CSV fields
    one(std::string(1 << 16, '\n'), "");
    auto run = [&](long count, uint64_t seed) {
        rng_state = seed;
        for (long i = 0; i < count; ++i) {
            std::string s = mutate(kCorpus[rnd() % n]);
            if ((rnd() & 3) == 0) s = mutate(s);
            one(s, (rnd() & 7) == 0 ? mutate("NVIDIA-SMI has failed") : "");
        }
    };
    // argv[2] = N: the same stream of inputs split over N threads calling the entry points at once (ThreadSanitizer
    // build: the host functions keep no shared mutable state — lazily built tables must be built race-free)
    const int threads = argc > 2 ? std::atoi(argv[2]) : 1;
    if (threads > 1) {
        std::vector<std::thread> th;
        for (int k = 0; k < threads; ++k) th.emplace_back(run, iters / threads, 0x9E3779B97F4A7C15ull + (uint64_t)k * 0x1234567ull);
        for (auto& x : th) x.join();
    } else {
        run(iters, rng_state);
    }
    std::printf("host fuzz ok: %ld inputs, sink %zu\n", iters, sink);
    return 0;
}
