// probe.hpp — the long-lived probe context behind the C ABI.
//
// Takes the slot of utils.RunNvidiaSmi + utils.CheckGPUVisible in
// handleAttachingState (internal/controller/composableresource_controller.go:259,275;
// internal/utils/gpus.go:666-689, 54-86).  One Device per managed GPU holds the
// resident sweep buffers (2*S bytes: pattern region + copy destination), a
// stream, events and the reduction scratch, so a warm probe is pure kernel time.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/croprobe.h"
#include "kernels.cuh"

namespace cro {

struct Device {
    int ordinal = -1;              // CUDA ordinal
    int index = -1;                // rank: position in the minor-sorted list
    cro_dev_info info{};
    std::mutex mu;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> evpool;   // per-sweep timing events of the full probe
    unsigned char* region = nullptr;   // [0,S) pattern, [S,2S) copy destination
    uint64_t sweep_bytes = 0;
    uint64_t seed = 0;
    bool filled = false;
    KernelPlan plan{};
    SweepScratch scratch{};
    SweepOut* d_out = nullptr;         // device sweep-result slots
    SweepOut* h_out = nullptr;         // pinned host mirror
    cro_probe_result* d_result = nullptr;  // all-gather send buffer
    cro_probe_result* d_gather = nullptr;  // all-gather receive buffer (world entries)
    unsigned long long* d_chase_next = nullptr;  // latency permutation (peers read it)
    unsigned long long* d_chase_out = nullptr;
    // asynchronous probe (ctx_probe_begin / ctx_probe_end)
    bool pending = false, have_pending_result = false;
    int pending_rc = 0;
    size_t pending_events = 0;
    std::chrono::steady_clock::time_point pending_since{};
    cro_probe_result pending_result{};
    // the whole probe captured as one CUDA graph (timing events are external event-record nodes)
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_key = 0;
    size_t graph_events = 0;
    bool graph_failed = false;
    bool have_expected = false;
    uint64_t expect_x = 0, expect_s = 0;
    unsigned sm_clock_mhz = 0, mem_clock_mhz = 0;
    uint32_t ecc_uncorrected = 0;      // NVML count cached at init / full-box probe / failed probe

    Device() = default;
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
    // Releases every CUDA object this device owns (probe.cu).  Runs for half-built devices too,
    // so a cro_probe_init that fails midway (OOM on the sweep region) leaks nothing.
    ~Device();
};

}  // namespace cro

struct cro_ctx {
    cro_opts opts{};
    std::vector<std::unique_ptr<cro::Device>> devs;   // minor-sorted
    std::atomic<uint64_t> launches{0};
    std::mutex err_mu;
    std::string last_error;
    std::mutex all_mu;                 // serialises cro_probe_all
    void* nccl_lib = nullptr;
    std::vector<void*> nccl_comms;     // ncclComm_t per device
    bool nccl_ready = false;
    bool peers_enabled = false;

    void set_error(const std::string& m) {
        std::lock_guard<std::mutex> g(err_mu);
        last_error = m;
    }
    // a context that dies during cro_probe_init hands its error text to the calling thread
    // (cro_last_error(NULL, ...)); defined in probe.cu
    ~cro_ctx();
};

namespace cro {

const std::string& last_init_error();   // calling thread's last failed ctx_create (or exception stopped at the C ABI)
void set_thread_error(const std::string& m) noexcept;
int ctx_create(const cro_opts* o, cro_ctx** out);
void ctx_destroy(cro_ctx* c);
int ctx_probe_device(cro_ctx* c, int idx, cro_probe_result* out);
int ctx_probe_all(cro_ctx* c, cro_probe_result* out, int cap, int* n);
int ctx_probe_begin(cro_ctx* c, int idx);
int ctx_probe_end(cro_ctx* c, int idx, cro_probe_result* out);
int ctx_probe_poll(cro_ctx* c, int idx);
int ctx_probe_wait(cro_ctx* c, int idx);

// single sweeps (each takes the device mutex)
int ctx_fill(cro_ctx* c, int idx, uint32_t iters, cro_sweep_result* out);
int ctx_read(cro_ctx* c, int idx, uint32_t variant, uint32_t iters, bool dst_half, cro_sweep_result* out);
int ctx_copy(cro_ctx* c, int idx, uint32_t variant, uint32_t iters, cro_sweep_result* out);
int ctx_expected(cro_ctx* c, int idx, cro_sweep_result* out);
int ctx_inject(cro_ctx* c, int idx, uint64_t word, uint64_t mask);
int ctx_read_words(cro_ctx* c, int idx, uint64_t first, uint64_t n, uint64_t* out);

uint32_t resolve_read_variant(uint32_t v, uint64_t bytes);
uint32_t resolve_copy_variant(uint32_t v);

}  // namespace cro
