// probe.cu — probe context: enumeration, resident sweep buffers, timed sweeps,
// concurrent multi-device probe with NVLink rounds and the one NCCL all-gather.
//
// Reference slot: utils.RunNvidiaSmi (internal/utils/gpus.go:666-689) and
// utils.CheckGPUVisible (internal/utils/gpus.go:54-86) as called from
// handleAttachingState (internal/controller/composableresource_controller.go:259,275).
#include "probe.hpp"

#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>

#include "identity.hpp"

namespace cro {

namespace {

constexpr int kMaxSweeps = 64;
constexpr uint64_t kDefaultSweep = 4ull << 30;
constexpr uint64_t kDefaultP2P = 1ull << 30;
constexpr uint64_t kDefaultSeedBase = 0x00C0FFEE00000000ull;
constexpr uint32_t kDefaultHops = 4096;
constexpr uint32_t kChaseSlots = 16384;      // one 8-byte slot per 128-byte line → 2 MiB

#define CU_TRY(ctx, expr)                                                              \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));     \
            return e__ == cudaErrorMemoryAllocation ? CRO_ERR_OOM : CRO_ERR_CUDA;      \
        }                                                                              \
    } while (0)

uint64_t ms_to_ns(float ms) { return (uint64_t)((double)ms * 1.0e6 + 0.5); }

int env_u32(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int ensure_region(cro_ctx* c, Device* d) {
    if (d->region) return CRO_OK;
    CU_TRY(c, cudaSetDevice(d->ordinal));
    // A device that is already in use may not have 2*S free (the reference's own pre-check for that
    // is CheckNoGPULoads, internal/utils/gpus.go:88).  Degrade: halve S down to 64 MiB — still far
    // beyond the 126 MB L2 when doubled — and report the size actually swept in the result.
    const uint64_t asked = d->sweep_bytes;
    cudaError_t e = cudaErrorMemoryAllocation;
    for (uint64_t s = asked;; s = (s / 2) & ~(uint64_t)15) {
        e = cudaMalloc(&d->region, 2 * s);
        if (e == cudaSuccess) {
            if (s != d->sweep_bytes) {
                d->sweep_bytes = s;
                d->have_expected = false;   // the closed form depends on S
                if (d->graph_exec) { cudaGraphExecDestroy(d->graph_exec); d->graph_exec = nullptr; }
            }
            break;
        }
        cudaGetLastError();
        d->region = nullptr;
        if (e != cudaErrorMemoryAllocation || s <= (64ull << 20) || !(c->opts.flags & CRO_F_DEGRADE_ON_OOM)) {
            c->set_error("cudaMalloc of sweep region (" + std::to_string(2 * s) + " bytes, asked for " +
                         std::to_string(2 * asked) + ") failed: " + cudaGetErrorString(e));
            return CRO_ERR_OOM;
        }
    }
    d->filled = false;
    return CRO_OK;
}

int ensure_filled(cro_ctx* c, Device* d) {
    int rc = ensure_region(c, d);
    if (rc) return rc;
    if (d->filled) return CRO_OK;
    CU_TRY(c, launch_fill(d->plan, d->region, d->sweep_bytes, d->seed, d->stream));
    c->launches++;
    d->filled = true;
    return CRO_OK;
}

int ensure_expected(cro_ctx* c, Device* d) {
    if (d->have_expected) return CRO_OK;
    CU_TRY(c, launch_expected(d->plan, d->sweep_bytes, d->seed, d->scratch, &d->d_out[kMaxSweeps - 1],
                              d->stream));
    c->launches++;
    CU_TRY(c, cudaMemcpyAsync(&d->h_out[kMaxSweeps - 1], &d->d_out[kMaxSweeps - 1], sizeof(SweepOut),
                              cudaMemcpyDeviceToHost, d->stream));
    CU_TRY(c, cudaStreamSynchronize(d->stream));
    d->expect_x = d->h_out[kMaxSweeps - 1].x;
    d->expect_s = d->h_out[kMaxSweeps - 1].s;
    d->have_expected = true;
    return CRO_OK;
}

// Waits for the stream, honouring opts.deadline_ms (kernels cannot be
// cancelled; on expiry the caller gets CRO_ERR_DEADLINE and the next call on
// this device synchronises first because it takes the same stream).
int wait_stream(cro_ctx* c, Device* d) {
    if (c->opts.deadline_ms <= 0) {
        CU_TRY(c, cudaStreamSynchronize(d->stream));
        return CRO_OK;
    }
    const auto until = std::chrono::steady_clock::now() + std::chrono::milliseconds(c->opts.deadline_ms);
    for (;;) {
        cudaError_t q = cudaStreamQuery(d->stream);
        if (q == cudaSuccess) return CRO_OK;
        if (q != cudaErrorNotReady) {
            c->set_error(std::string("cudaStreamQuery: ") + cudaGetErrorString(q));
            return CRO_ERR_CUDA;
        }
        if (std::chrono::steady_clock::now() > until) {
            c->set_error("probe deadline of " + std::to_string(c->opts.deadline_ms) + " ms exceeded");
            return CRO_ERR_DEADLINE;
        }
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
}

void copy_cstr(char* dst, size_t cap, const std::string& s) {
    memset(dst, 0, cap);
    memcpy(dst, s.data(), std::min(cap - 1, s.size()));
}

}  // namespace

uint32_t resolve_read_variant(uint32_t v, uint64_t bytes) {
    // AUTO: the TMA ring wins from ~1 GiB up (7.46 vs 7.31 TB/s at 4 GiB); for small sweeps its fixed
    // cost (one CTA per SM, atomic tile claims) loses to plain LDG (profiles/r01_size_sweep.jsonl).
    if (v == CRO_READ_AUTO) v = (uint32_t)env_u32("CRO_READ_VARIANT", bytes <= (128ull << 20) ? CRO_READ_LDG : CRO_READ_TMA);
    return (v == READ_LDG || v == READ_TMA || v == READ_LDG256) ? v : (uint32_t)READ_TMA;
}
uint32_t resolve_copy_variant(uint32_t v) {
    if (v == CRO_COPY_AUTO) v = (uint32_t)env_u32("CRO_COPY_VARIANT", CRO_COPY_TMA);
    return (v == COPY_LDG || v == COPY_TMA) ? v : (uint32_t)COPY_TMA;
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
// Caches the device's uncorrected volatile ECC count (0 when NVML is not the identity source or ECC is off).
static void refresh_ecc(cro_ctx* c, Device* d) {
    if ((c->opts.flags & CRO_F_NO_NVML) || d->info.identity_source != 1) return;
    unsigned long long ecc = 0;
    if (identity::NvmlEccUncorrected(std::string(d->info.gpu_uuid, strnlen(d->info.gpu_uuid, sizeof d->info.gpu_uuid)), &ecc))
        d->ecc_uncorrected = (uint32_t)std::min<unsigned long long>(ecc, 0xFFFFFFFFull);
}

int ctx_create(const cro_opts* o, cro_ctx** out) {
    if (!out) return CRO_ERR_INVALID_ARG;
    *out = nullptr;
    cro_opts opts;
    memset(&opts, 0, sizeof opts);
    if (o) opts = *o;
    else opts.abi_version = CRO_ABI_VERSION;
    if (opts.abi_version != CRO_ABI_VERSION) return CRO_ERR_ABI_MISMATCH;
    if (opts.sweep_bytes == 0) opts.sweep_bytes = kDefaultSweep;
    if (opts.sweep_bytes % 16 != 0 || opts.sweep_bytes < 16) return CRO_ERR_INVALID_ARG;
    if (opts.p2p_bytes == 0) opts.p2p_bytes = std::min(kDefaultP2P, opts.sweep_bytes);
    if (opts.p2p_bytes > opts.sweep_bytes || opts.p2p_bytes % 16 != 0) return CRO_ERR_INVALID_ARG;
    if (opts.seed_base == 0) opts.seed_base = kDefaultSeedBase;
    if (opts.read_sweeps == 0) opts.read_sweeps = 5;
    if (opts.copy_sweeps == 0) opts.copy_sweeps = 5;
    if (opts.read_sweeps > kMaxSweeps - 2 || opts.copy_sweeps > kMaxSweeps - 2) return CRO_ERR_INVALID_ARG;
    if (opts.latency_hops == 0) opts.latency_hops = kDefaultHops;
    if (opts.n_devices < 0 || opts.n_devices > CRO_MAX_DEVICES) return CRO_ERR_INVALID_ARG;

    std::unique_ptr<cro_ctx> c(new cro_ctx);
    c->opts = opts;

    int n_cuda = 0;
    cudaError_t e = cudaGetDeviceCount(&n_cuda);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
        // No usable GPU.  A probe library without a GPU must say so loudly:
        // there is no CPU fallback on this path.
        cudaGetLastError();
        return CRO_ERR_NO_DEVICE;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return CRO_ERR_CUDA;
    }
    std::vector<int> ordinals;
    if (opts.n_devices > 0) {
        for (int i = 0; i < opts.n_devices; ++i) {
            if (opts.devices[i] < 0 || opts.devices[i] >= n_cuda) return CRO_ERR_INVALID_ARG;
            ordinals.push_back(opts.devices[i]);
        }
    } else {
        for (int i = 0; i < n_cuda && i < CRO_MAX_DEVICES; ++i) ordinals.push_back(i);
    }

    std::vector<identity::NvmlGpu> nvml;
    bool have_nvml = false;
    if (!(opts.flags & CRO_F_NO_NVML)) have_nvml = identity::ScanNvml(&nvml, nullptr);
    const std::vector<identity::ProcGpu> proc = identity::ScanProc("/proc");

    struct Keyed { std::unique_ptr<Device> d; long long key; };
    std::vector<Keyed> keyed;
    for (int ord : ordinals) {
        std::unique_ptr<Device> d(new Device);
        d->ordinal = ord;
        cudaDeviceProp prop;
        CU_TRY(c.get(), cudaGetDeviceProperties(&prop, ord));
        cro_dev_info& info = d->info;
        memset(&info, 0, sizeof info);
        info.cuda_ordinal = ord;
        info.device_minor = -1;
        const std::string uuid = identity::FormatGpuUuid(reinterpret_cast<const unsigned char*>(prop.uuid.bytes));
        copy_cstr(info.gpu_uuid, sizeof info.gpu_uuid, uuid);
        copy_cstr(info.pci_bus_id, sizeof info.pci_bus_id,
                  identity::FormatBusIdSmi((unsigned)prop.pciDomainID, (unsigned)prop.pciBusID,
                                           (unsigned)prop.pciDeviceID, 0));
        copy_cstr(info.name, sizeof info.name, prop.name);
        info.hbm_bytes_total = prop.totalGlobalMem;
        info.sm_count = (uint32_t)prop.multiProcessorCount;
        info.cc_major = (uint32_t)prop.major;
        info.cc_minor = (uint32_t)prop.minor;
        info.identity_source = 3;
        long long key = ((long long)prop.pciDomainID << 16) | ((long long)prop.pciBusID << 8) |
                        (long long)prop.pciDeviceID;
        bool matched = false;
        if (have_nvml) {
            for (size_t k = 0; k < nvml.size(); ++k) {
                if (nvml[k].uuid != uuid) continue;
                info.device_minor = nvml[k].minor;
                if (!nvml[k].bus_id.empty()) copy_cstr(info.pci_bus_id, sizeof info.pci_bus_id, nvml[k].bus_id);
                info.identity_source = 1;
                d->sm_clock_mhz = nvml[k].sm_clock_mhz;
                d->mem_clock_mhz = nvml[k].mem_clock_mhz;
                key = (long long)k;   // nvidia-smi lists in NVML index order
                matched = true;
                break;
            }
        }
        if (!matched) {
            for (const identity::ProcGpu& g : proc) {
                if (g.uuid != uuid) continue;
                info.device_minor = atoi(g.minor.c_str());
                info.identity_source = 2;
                break;
            }
        }
        keyed.push_back({std::move(d), key});
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const Keyed& a, const Keyed& b) { return a.key < b.key; });

    for (size_t i = 0; i < keyed.size(); ++i) {
        Device* d = keyed[i].d.get();
        d->index = (int)i;
        d->sweep_bytes = opts.sweep_bytes;
        d->seed = opts.seed_base | (uint64_t)(d->info.device_minor >= 0 ? d->info.device_minor : d->ordinal);
        CU_TRY(c.get(), cudaSetDevice(d->ordinal));
        CU_TRY(c.get(), cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
        CU_TRY(c.get(), cudaEventCreate(&d->ev0));
        CU_TRY(c.get(), cudaEventCreate(&d->ev1));
        CU_TRY(c.get(), plan_kernels(d->ordinal, &d->plan));
        int max_grid = std::max({d->plan.fill.grid, d->plan.read_ldg.grid, d->plan.read_ldg256.grid,
                                 d->plan.read_tma.grid, d->plan.expect.grid, 1});
        CU_TRY(c.get(), cudaMalloc(&d->scratch.partials, sizeof(ulonglong2) * (size_t)max_grid));
        CU_TRY(c.get(), cudaMalloc(&d->scratch.counter, sizeof(unsigned)));
        CU_TRY(c.get(), cudaMalloc(&d->scratch.tmin, sizeof(unsigned long long)));
        CU_TRY(c.get(), cudaMalloc(&d->scratch.tmax, sizeof(unsigned long long)));
        CU_TRY(c.get(), cudaMalloc(&d->scratch.tile_ctr, sizeof(unsigned long long)));
        CU_TRY(c.get(), cudaMemset(d->scratch.tile_ctr, 0, sizeof(unsigned long long)));
        CU_TRY(c.get(), cudaMemset(d->scratch.counter, 0, sizeof(unsigned)));
        CU_TRY(c.get(), cudaMemset(d->scratch.tmin, 0xFF, sizeof(unsigned long long)));
        CU_TRY(c.get(), cudaMemset(d->scratch.tmax, 0, sizeof(unsigned long long)));
        CU_TRY(c.get(), cudaMalloc(&d->d_out, sizeof(SweepOut) * kMaxSweeps));
        CU_TRY(c.get(), cudaMallocHost(&d->h_out, sizeof(SweepOut) * kMaxSweeps));
        CU_TRY(c.get(), cudaMalloc(&d->d_result, sizeof(cro_probe_result)));
        CU_TRY(c.get(), cudaMalloc(&d->d_gather, sizeof(cro_probe_result) * CRO_MAX_DEVICES));
        CU_TRY(c.get(), cudaMemset(d->d_result, 0, sizeof(cro_probe_result)));
        CU_TRY(c.get(), cudaMalloc(&d->d_chase_out, 4 * sizeof(unsigned long long)));
        if (!(opts.flags & CRO_F_LAZY_ALLOC)) {
            int rc = ensure_region(c.get(), d);
            if (rc) return rc;
        }
        refresh_ecc(c.get(), d);
        c->devs.push_back(std::move(keyed[i].d));
    }
    *out = c.release();
    return CRO_OK;
}

thread_local std::string g_init_error;
const std::string& last_init_error() { return g_init_error; }
void set_thread_error(const std::string& m) noexcept {
    try { g_init_error = m; } catch (...) {}
}

}  // namespace cro
cro_ctx::~cro_ctx() {
    if (!last_error.empty()) cro::g_init_error = last_error;
}
namespace cro {

Device::~Device() {
    if (ordinal < 0) return;                      // never bound to a CUDA device: owns nothing
    cudaSetDevice(ordinal);
    if (stream) cudaStreamSynchronize(stream);
    cudaFree(region);                             // cudaFree(nullptr) is a no-op
    cudaFree(scratch.partials);
    cudaFree(scratch.counter);
    cudaFree(scratch.tmin);
    cudaFree(scratch.tmax);
    cudaFree(scratch.tile_ctr);
    cudaFree(d_out);
    if (h_out) cudaFreeHost(h_out);
    cudaFree(d_result);
    cudaFree(d_gather);
    cudaFree(d_chase_next);
    cudaFree(d_chase_out);
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (cudaEvent_t e : evpool) cudaEventDestroy(e);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    cudaGetLastError();                           // a failed release must not poison the caller's next CUDA call
}

void ctx_destroy(cro_ctx* c) {
    if (!c) return;
    if (c->nccl_ready && c->nccl_lib) {
        auto destroy = (int (*)(void*))dlsym(c->nccl_lib, "ncclCommDestroy");
        if (destroy)
            for (void* comm : c->nccl_comms)
                if (comm) destroy(comm);
    }
    delete c;                                     // ~Device releases the per-device CUDA objects
}

static void drain_pending_fwd(cro_ctx* c, Device* d);

static Device* dev_at(cro_ctx* c, int idx) {
    if (!c || idx < 0 || idx >= (int)c->devs.size()) return nullptr;
    return c->devs[(size_t)idx].get();
}

// ---------------------------------------------------------------------------
// single sweeps
// ---------------------------------------------------------------------------
int ctx_fill(cro_ctx* c, int idx, uint32_t iters, cro_sweep_result* out) {
    Device* d = dev_at(c, idx);
    if (!d || !out || iters == 0) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    drain_pending_fwd(c, d);
    CU_TRY(c, cudaSetDevice(d->ordinal));
    int rc = ensure_region(c, d);
    if (rc) return rc;
    CU_TRY(c, cudaEventRecord(d->ev0, d->stream));
    for (uint32_t i = 0; i < iters; ++i)
        CU_TRY(c, launch_fill(d->plan, d->region, d->sweep_bytes, d->seed, d->stream));
    CU_TRY(c, cudaEventRecord(d->ev1, d->stream));
    c->launches += iters;
    d->filled = true;
    if ((rc = wait_stream(c, d))) return rc;
    float ms = 0;
    CU_TRY(c, cudaEventElapsedTime(&ms, d->ev0, d->ev1));
    memset(out, 0, sizeof *out);
    out->bytes = d->sweep_bytes * iters;
    out->ns = ms_to_ns(ms);
    out->launches = iters;
    return CRO_OK;
}

int ctx_read(cro_ctx* c, int idx, uint32_t variant, uint32_t iters, bool dst_half,
             cro_sweep_result* out) {
    Device* d = dev_at(c, idx);
    if (!d || !out || iters == 0) return CRO_ERR_INVALID_ARG;
    variant = resolve_read_variant(variant, d->sweep_bytes);
    std::lock_guard<std::mutex> g(d->mu);
    drain_pending_fwd(c, d);
    CU_TRY(c, cudaSetDevice(d->ordinal));
    int rc = ensure_filled(c, d);
    if (rc) return rc;
    const unsigned char* base = d->region + (dst_half ? d->sweep_bytes : 0);
    CU_TRY(c, cudaEventRecord(d->ev0, d->stream));
    for (uint32_t i = 0; i < iters; ++i)
        CU_TRY(c, launch_read(d->plan, variant, base, d->sweep_bytes, d->scratch, &d->d_out[0], d->stream));
    CU_TRY(c, cudaEventRecord(d->ev1, d->stream));
    c->launches += iters;
    CU_TRY(c, cudaMemcpyAsync(&d->h_out[0], &d->d_out[0], sizeof(SweepOut), cudaMemcpyDeviceToHost,
                              d->stream));
    if ((rc = wait_stream(c, d))) return rc;
    float ms = 0;
    CU_TRY(c, cudaEventElapsedTime(&ms, d->ev0, d->ev1));
    memset(out, 0, sizeof *out);
    out->bytes = d->sweep_bytes * iters;
    out->ns = ms_to_ns(ms);
    out->checksum_xor = d->h_out[0].x;
    out->checksum_sum = d->h_out[0].s;
    out->variant = variant;
    out->launches = iters;
    return CRO_OK;
}

int ctx_copy(cro_ctx* c, int idx, uint32_t variant, uint32_t iters, cro_sweep_result* out) {
    Device* d = dev_at(c, idx);
    if (!d || !out || iters == 0) return CRO_ERR_INVALID_ARG;
    variant = resolve_copy_variant(variant);
    std::lock_guard<std::mutex> g(d->mu);
    drain_pending_fwd(c, d);
    CU_TRY(c, cudaSetDevice(d->ordinal));
    int rc = ensure_filled(c, d);
    if (rc) return rc;
    CU_TRY(c, cudaEventRecord(d->ev0, d->stream));
    for (uint32_t i = 0; i < iters; ++i)
        CU_TRY(c, launch_copy(d->plan, variant, d->region + d->sweep_bytes, d->region, d->sweep_bytes,
                              d->scratch, d->stream));
    CU_TRY(c, cudaEventRecord(d->ev1, d->stream));
    c->launches += iters;
    if ((rc = wait_stream(c, d))) return rc;
    float ms = 0;
    CU_TRY(c, cudaEventElapsedTime(&ms, d->ev0, d->ev1));
    memset(out, 0, sizeof *out);
    out->bytes = 2 * d->sweep_bytes * iters;
    out->ns = ms_to_ns(ms);
    out->variant = variant;
    out->launches = iters;
    return CRO_OK;
}

int ctx_expected(cro_ctx* c, int idx, cro_sweep_result* out) {
    Device* d = dev_at(c, idx);
    if (!d || !out) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    drain_pending_fwd(c, d);
    CU_TRY(c, cudaSetDevice(d->ordinal));
    d->have_expected = false;
    CU_TRY(c, cudaEventRecord(d->ev0, d->stream));
    int rc = ensure_expected(c, d);
    if (rc) return rc;
    CU_TRY(c, cudaEventRecord(d->ev1, d->stream));
    CU_TRY(c, cudaStreamSynchronize(d->stream));
    float ms = 0;
    CU_TRY(c, cudaEventElapsedTime(&ms, d->ev0, d->ev1));
    memset(out, 0, sizeof *out);
    out->bytes = 0;
    out->ns = ms_to_ns(ms);
    out->checksum_xor = d->expect_x;
    out->checksum_sum = d->expect_s;
    out->launches = 1;
    return CRO_OK;
}

int ctx_inject(cro_ctx* c, int idx, uint64_t word, uint64_t mask) {
    Device* d = dev_at(c, idx);
    if (!d) return CRO_ERR_INVALID_ARG;
    if (word >= d->sweep_bytes / 8) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    drain_pending_fwd(c, d);
    CU_TRY(c, cudaSetDevice(d->ordinal));
    int rc = ensure_filled(c, d);
    if (rc) return rc;
    CU_TRY(c, launch_xor_word(d->region, word, mask, d->stream));
    c->launches++;
    CU_TRY(c, cudaStreamSynchronize(d->stream));
    return CRO_OK;
}

int ctx_read_words(cro_ctx* c, int idx, uint64_t first, uint64_t n, uint64_t* out) {
    Device* d = dev_at(c, idx);
    if (!d || !out) return CRO_ERR_INVALID_ARG;
    if (first + n > 2 * d->sweep_bytes / 8) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    drain_pending_fwd(c, d);
    CU_TRY(c, cudaSetDevice(d->ordinal));
    int rc = ensure_filled(c, d);
    if (rc) return rc;
    CU_TRY(c, cudaMemcpyAsync(out, d->region + first * 8, n * 8, cudaMemcpyDeviceToHost, d->stream));
    CU_TRY(c, cudaStreamSynchronize(d->stream));
    return CRO_OK;
}

// ---------------------------------------------------------------------------
// full per-device probe
// ---------------------------------------------------------------------------
static uint64_t median_of(std::vector<uint64_t> v) {
    std::sort(v.begin(), v.end());
    return v.empty() ? 0 : v[v.size() / 2];
}

// Caller holds d->mu.  Enqueues one whole probe (fill, reads, copies, result
// copy-back) on the device's stream and returns without waiting.
static int probe_enqueue(cro_ctx* c, Device* d, cro_probe_result* r) {
    const cro_opts& o = c->opts;
    memset(r, 0, sizeof *r);
    r->abi_version = CRO_ABI_VERSION;
    r->cuda_ordinal = d->ordinal;
    r->device_minor = d->info.device_minor;
    memcpy(r->gpu_uuid, d->info.gpu_uuid, sizeof r->gpu_uuid);
    memcpy(r->pci_bus_id, d->info.pci_bus_id, sizeof r->pci_bus_id);
    r->hbm_bytes_total = d->info.hbm_bytes_total;
    r->sweep_bytes = d->sweep_bytes;
    r->seed = d->seed;
    r->sm_count = d->info.sm_count;
    r->sm_clock_mhz = d->sm_clock_mhz;
    r->mem_clock_mhz = d->mem_clock_mhz;
    r->rank = c->opts.rank_base + (uint32_t)d->index;
    r->world = c->opts.world_override ? c->opts.world_override : (uint32_t)c->devs.size();
    r->p2p_bytes = o.p2p_bytes;
    const uint32_t rv = resolve_read_variant(o.read_variant, d->sweep_bytes);
    const uint32_t cv = resolve_copy_variant(o.copy_variant);
    r->read_variant = rv;
    r->copy_variant = cv;
    r->read_sweeps = o.read_sweeps;
    r->copy_sweeps = (o.flags & CRO_F_SKIP_COPY) ? 0 : o.copy_sweeps;

    CU_TRY(c, cudaSetDevice(d->ordinal));
    int rc = ensure_region(c, d);
    if (rc) { r->status = rc; return rc; }
    if ((rc = ensure_expected(c, d))) { r->status = rc; return rc; }
    r->sweep_bytes = d->sweep_bytes;   // ensure_region may have degraded it (CRO_F_DEGRADE_ON_OOM)
    r->expect_xor = d->expect_x;
    r->expect_sum = d->expect_s;

    // events: fill | R reads | C copies   (pool lives with the device)
    const size_t need = 2 + o.read_sweeps + r->copy_sweeps + 2;
    while (d->evpool.size() < need) {
        cudaEvent_t e;
        CU_TRY(c, cudaEventCreate(&e));
        d->evpool.push_back(e);
    }
    std::vector<cudaEvent_t>& ev = d->evpool;

    const bool verify = (o.flags & CRO_F_VERIFY_COPY) && r->copy_sweeps > 0;
    size_t k = 0;
    // The whole probe as one sequence; `external` records the timing events as external event-record
    // nodes so that the same sequence can be stream-captured into a CUDA graph once and replayed.
    auto issue = [&](bool external) -> int {
        const unsigned flag = external ? cudaEventRecordExternal : cudaEventRecordDefault;
        k = 0;
        CU_TRY(c, cudaEventRecordWithFlags(ev[k++], d->stream, flag));
        CU_TRY(c, launch_fill(d->plan, d->region, d->sweep_bytes, d->seed, d->stream));
        CU_TRY(c, cudaEventRecordWithFlags(ev[k++], d->stream, flag));
        for (uint32_t i = 0; i < o.read_sweeps; ++i) {
            CU_TRY(c, launch_read(d->plan, rv, d->region, d->sweep_bytes, d->scratch, &d->d_out[i], d->stream));
            CU_TRY(c, cudaEventRecordWithFlags(ev[k++], d->stream, flag));
        }
        for (uint32_t i = 0; i < r->copy_sweeps; ++i) {
            CU_TRY(c, launch_copy(d->plan, cv, d->region + d->sweep_bytes, d->region, d->sweep_bytes, d->scratch, d->stream));
            CU_TRY(c, cudaEventRecordWithFlags(ev[k++], d->stream, flag));
        }
        if (verify)
            CU_TRY(c, launch_read(d->plan, rv, d->region + d->sweep_bytes, d->sweep_bytes, d->scratch,
                                  &d->d_out[o.read_sweeps], d->stream));
        CU_TRY(c, cudaMemcpyAsync(d->h_out, d->d_out, sizeof(SweepOut) * (o.read_sweeps + 1),
                                  cudaMemcpyDeviceToHost, d->stream));
        return CRO_OK;
    };
    // One graph launch instead of ~35 runtime calls per probe (matters when one host thread feeds 8 GPUs).
    // The graph is tied to the options it was captured with; any capture problem falls back to direct launches.
    const uint64_t graph_key = ((uint64_t)rv << 48) ^ ((uint64_t)cv << 40) ^ ((uint64_t)o.read_sweeps << 24) ^
                               ((uint64_t)r->copy_sweeps << 8) ^ (verify ? 1u : 0u);
    if (env_u32("CRO_USE_GRAPH", 1) && !d->graph_failed) {
        if (d->graph_exec && d->graph_key != graph_key) {
            cudaGraphExecDestroy(d->graph_exec);
            d->graph_exec = nullptr;
        }
        if (!d->graph_exec) {
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(d->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                const int irc = issue(true);
                const cudaError_t ec = cudaStreamEndCapture(d->stream, &graph);
                ok = irc == CRO_OK && ec == cudaSuccess && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&d->graph_exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            if (!ok) {
                cudaGetLastError();
                d->graph_exec = nullptr;
                d->graph_failed = true;
            } else {
                d->graph_key = graph_key;
                d->graph_events = k;
            }
        }
    }
    if (d->graph_exec) {
        CU_TRY(c, cudaGraphLaunch(d->graph_exec, d->stream));
        k = d->graph_events;
    } else {
        int irc = issue(false);
        if (irc) return irc;
    }
    d->filled = true;
    c->launches += 1 + o.read_sweeps + r->copy_sweeps + (verify ? 1 : 0);
    d->pending_events = k;
    return CRO_OK;
}

// Caller holds d->mu.  Waits for the probe enqueued by probe_enqueue and
// evaluates it into *r (which probe_enqueue started filling).
static int probe_finish(cro_ctx* c, Device* d, cro_probe_result* r) {
    const cro_opts& o = c->opts;
    std::vector<cudaEvent_t>& ev = d->evpool;
    const size_t k = d->pending_events;
    const bool verify = (o.flags & CRO_F_VERIFY_COPY) && r->copy_sweeps > 0;
    CU_TRY(c, cudaSetDevice(d->ordinal));
    int rc;
    if ((rc = wait_stream(c, d))) { r->status = rc; return rc; }

    float ms = 0;
    CU_TRY(c, cudaEventElapsedTime(&ms, ev[0], ev[1]));
    r->fill_ns = ms_to_ns(ms);
    std::vector<uint64_t> rt, ct;
    for (uint32_t i = 0; i < o.read_sweeps; ++i) {
        CU_TRY(c, cudaEventElapsedTime(&ms, ev[1 + i], ev[2 + i]));
        rt.push_back(ms_to_ns(ms));
    }
    for (uint32_t i = 0; i < r->copy_sweeps; ++i) {
        CU_TRY(c, cudaEventElapsedTime(&ms, ev[1 + o.read_sweeps + i], ev[2 + o.read_sweeps + i]));
        ct.push_back(ms_to_ns(ms));
    }
    CU_TRY(c, cudaEventElapsedTime(&ms, ev[0], ev[k - 1]));
    r->total_ns = ms_to_ns(ms);
    for (uint64_t t : rt) r->read_total_ns += t;
    for (uint64_t t : ct) r->copy_total_ns += t;
    r->read_best_ns = *std::min_element(rt.begin(), rt.end());
    r->read_median_ns = median_of(rt);
    if (!ct.empty()) {
        r->copy_best_ns = *std::min_element(ct.begin(), ct.end());
        r->copy_median_ns = median_of(ct);
    }
    // every sweep must reproduce the closed form, not just the last one
    r->checksum_xor = d->h_out[0].x;
    r->checksum_sum = d->h_out[0].s;
    int status = CRO_OK;
    for (uint32_t i = 0; i < o.read_sweeps; ++i) {
        if (d->h_out[i].x != d->expect_x || d->h_out[i].s != d->expect_s) {
            r->checksum_xor = d->h_out[i].x;
            r->checksum_sum = d->h_out[i].s;
            status = CRO_ERR_CHECKSUM;
            c->set_error("HBM read sweep " + std::to_string(i) + " on " + d->info.gpu_uuid +
                         " does not reproduce the pattern checksum");
            break;
        }
    }
    if (verify) {
        r->copy_checksum_xor = d->h_out[o.read_sweeps].x;
        r->copy_checksum_sum = d->h_out[o.read_sweeps].s;
        if (status == CRO_OK && (r->copy_checksum_xor != d->expect_x || r->copy_checksum_sum != d->expect_s)) {
            status = CRO_ERR_CHECKSUM;
            c->set_error(std::string("HBM copy destination on ") + d->info.gpu_uuid +
                         " does not reproduce the pattern checksum");
        }
    }
    // What the memory itself reported: uncorrected volatile ECC errors (nvmlDeviceGetTotalEccErrors).
    // NVML calls serialise across processes (measured: ~2 ms each with 4 ranks probing, enough to skew the
    // ranks' all-gather), so the warm probe reuses the count read at init / at the last full-box probe and
    // only a FAILED probe pays for a fresh read.
    if (status != CRO_OK) refresh_ecc(c, d);
    r->ecc_errors = d->ecc_uncorrected;
    r->status = status;
    return status;
}

static int probe_locked(cro_ctx* c, Device* d, cro_probe_result* r) {
    int rc = probe_enqueue(c, d, r);
    if (rc) { r->status = rc; return rc; }
    return probe_finish(c, d, r);
}

// Drains a probe begun with ctx_probe_begin whose result nobody has collected
// yet, so another operation may use the stream / result slots.  Caller holds d->mu.
static void drain_pending(cro_ctx* c, Device* d) {
    if (!d->pending) return;
    d->pending_rc = probe_finish(c, d, &d->pending_result);
    d->pending = false;
    d->have_pending_result = true;
}

static void drain_pending_fwd(cro_ctx* c, Device* d) { drain_pending(c, d); }

static int publish_result(cro_ctx* c, Device* d, const cro_probe_result* r) {
    CU_TRY(c, cudaSetDevice(d->ordinal));
    CU_TRY(c, cudaMemcpyAsync(d->d_result, r, sizeof *r, cudaMemcpyHostToDevice, d->stream));
    CU_TRY(c, cudaStreamSynchronize(d->stream));
    return CRO_OK;
}

int ctx_probe_device(cro_ctx* c, int idx, cro_probe_result* out) {
    Device* d = dev_at(c, idx);
    if (!d || !out) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    drain_pending(c, d);
    d->have_pending_result = false;   // a synchronous probe supersedes an uncollected asynchronous one
    int rc = probe_locked(c, d, out);
    if (rc == CRO_OK || rc == CRO_ERR_CHECKSUM) {
        int prc = publish_result(c, d, out);
        if (prc) return prc;
    }
    return rc;
}

// Asynchronous form: begin enqueues the probe and returns; end waits and
// evaluates.  Lets ONE host thread (the reference's single reconcile worker)
// keep every attached GPU busy: probes of different devices overlap.
int ctx_probe_begin(cro_ctx* c, int idx) {
    Device* d = dev_at(c, idx);
    if (!d) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    if (d->pending || d->have_pending_result) return CRO_OK;   // one in flight (or waiting to be collected)
    int rc = probe_enqueue(c, d, &d->pending_result);
    if (rc) return rc;
    d->pending = true;
    d->pending_since = std::chrono::steady_clock::now();
    return CRO_OK;
}

// 1 when a probe begun on this device has finished (or none is in flight), 0 while it runs.
int ctx_probe_poll(cro_ctx* c, int idx) {
    Device* d = dev_at(c, idx);
    if (!d) return 1;
    std::lock_guard<std::mutex> g(d->mu);
    if (!d->pending) return 1;
    cudaSetDevice(d->ordinal);
    // anything but "still running" counts as finished: a failed stream must not keep a poller spinning —
    // cro_probe_end then reports the CUDA error
    return cudaStreamQuery(d->stream) == cudaErrorNotReady ? 0 : 1;
}

// Blocks until the probe in flight on this device (if any) has finished; does not collect it.
int ctx_probe_wait(cro_ctx* c, int idx) {
    Device* d = dev_at(c, idx);
    if (!d) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    if (!d->pending) return CRO_OK;
    CU_TRY(c, cudaSetDevice(d->ordinal));
    CU_TRY(c, cudaStreamSynchronize(d->stream));
    return CRO_OK;
}

int ctx_probe_end(cro_ctx* c, int idx, cro_probe_result* out) {
    Device* d = dev_at(c, idx);
    if (!d || !out) return CRO_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    // a prefetched result nobody collected for more than a second says nothing about the device NOW
    if (d->have_pending_result && !d->pending &&
        std::chrono::steady_clock::now() - d->pending_since > std::chrono::seconds(1))
        d->have_pending_result = false;
    if (!d->pending && !d->have_pending_result) {   // nothing begun: behave like the synchronous call
        int rc = probe_enqueue(c, d, &d->pending_result);
        if (rc) return rc;
        d->pending = true;
        d->pending_since = std::chrono::steady_clock::now();
    }
    drain_pending(c, d);
    d->have_pending_result = false;
    *out = d->pending_result;
    int rc = d->pending_rc;
    if (rc == CRO_OK || rc == CRO_ERR_CHECKSUM) {
        int prc = publish_result(c, d, out);
        if (prc) return prc;
    }
    return rc;
}

// ---------------------------------------------------------------------------
// multi-device: concurrent probes, NVLink rounds, one all-gather
// ---------------------------------------------------------------------------
namespace {

// Round-robin 1-factorisation of K_n (n even): n-1 rounds of n/2 disjoint pairs.
std::vector<std::vector<std::pair<int, int>>> one_factorisation(int n) {
    std::vector<std::vector<std::pair<int, int>>> rounds;
    if (n < 2) return rounds;
    const int m = (n % 2 == 0) ? n : n + 1;  // odd n: vertex m-1 is a bye
    for (int r = 0; r < m - 1; ++r) {
        std::vector<std::pair<int, int>> pairs;
        auto add = [&](int a, int b) { if (a < n && b < n) pairs.push_back({a, b}); };
        add(m - 1, r);
        for (int k = 1; k < m / 2; ++k) add((r + k) % (m - 1), (r - k + (m - 1)) % (m - 1));
        rounds.push_back(pairs);
    }
    return rounds;
}

int enable_peers(cro_ctx* c) {
    if (c->peers_enabled) return CRO_OK;
    const int n = (int)c->devs.size();
    for (int a = 0; a < n; ++a) {
        CU_TRY(c, cudaSetDevice(c->devs[a]->ordinal));
        for (int b = 0; b < n; ++b) {
            if (a == b) continue;
            int can = 0;
            CU_TRY(c, cudaDeviceCanAccessPeer(&can, c->devs[a]->ordinal, c->devs[b]->ordinal));
            if (!can) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(c->devs[b]->ordinal, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                c->set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                cudaGetLastError();
                return CRO_ERR_P2P;
            }
            cudaGetLastError();
        }
    }
    c->peers_enabled = true;
    return CRO_OK;
}

// Sattolo cycle over kChaseSlots slots, mt19937_64 seeded from the owner's
// minor; slot i lives at next[i*16] (one per 128-byte line).
int ensure_chase(cro_ctx* c, Device* d) {
    if (d->d_chase_next) return CRO_OK;
    std::vector<unsigned long long> perm(kChaseSlots);
    for (uint32_t i = 0; i < kChaseSlots; ++i) perm[i] = i;
    std::mt19937_64 rng(0x5A77011000000000ull + (uint64_t)(d->info.device_minor >= 0 ? d->info.device_minor : d->ordinal));
    for (uint32_t i = kChaseSlots - 1; i > 0; --i) {
        const uint32_t j = (uint32_t)(rng() % i);
        std::swap(perm[i], perm[j]);
    }
    std::vector<unsigned long long> lines((size_t)kChaseSlots * 16, 0);
    for (uint32_t i = 0; i < kChaseSlots; ++i) lines[(size_t)i * 16] = perm[i];
    CU_TRY(c, cudaSetDevice(d->ordinal));
    CU_TRY(c, cudaMalloc(&d->d_chase_next, lines.size() * sizeof(unsigned long long)));
    CU_TRY(c, cudaMemcpy(d->d_chase_next, lines.data(), lines.size() * sizeof(unsigned long long),
                         cudaMemcpyHostToDevice));
    return CRO_OK;
}

struct Nccl {
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

int load_nccl(cro_ctx* c, Nccl* n) {
    if (!c->nccl_lib) {
        c->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!c->nccl_lib) c->nccl_lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!c->nccl_lib) {
            // torch wheels carry their own copy
            const char* extra = getenv("CRO_NCCL_PATH");
            if (extra) c->nccl_lib = dlopen(extra, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!c->nccl_lib) {
            c->set_error("libnccl.so.2 not found (set CRO_NCCL_PATH)");
            return CRO_ERR_NCCL;
        }
    }
    n->CommInitAll = (int (*)(void**, int, const int*))dlsym(c->nccl_lib, "ncclCommInitAll");
    n->GroupStart = (int (*)())dlsym(c->nccl_lib, "ncclGroupStart");
    n->GroupEnd = (int (*)())dlsym(c->nccl_lib, "ncclGroupEnd");
    n->AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(c->nccl_lib, "ncclAllGather");
    n->GetErrorString = (const char* (*)(int))dlsym(c->nccl_lib, "ncclGetErrorString");
    if (!n->CommInitAll || !n->GroupStart || !n->GroupEnd || !n->AllGather) {
        c->set_error("libnccl lacks a required symbol");
        return CRO_ERR_NCCL;
    }
    return CRO_OK;
}

}  // namespace

int ctx_probe_all(cro_ctx* c, cro_probe_result* out, int cap, int* n_out) {
    if (!c || !out || !n_out) return CRO_ERR_INVALID_ARG;
    const int n = (int)c->devs.size();
    *n_out = n;
    if (cap < n) return CRO_ERR_BUFFER_SMALL;
    if (n == 0) return CRO_OK;
    std::lock_guard<std::mutex> all(c->all_mu);
    const cro_opts& o = c->opts;

    // phase 1: every device probes concurrently, one host thread + stream each
    std::vector<cro_probe_result> res((size_t)n);
    std::vector<int> rcs((size_t)n, CRO_OK);
    {
        std::vector<std::thread> th;
        for (int i = 0; i < n; ++i)
            th.emplace_back([&, i] {
                try {
                    Device* d = c->devs[(size_t)i].get();
                    std::lock_guard<std::mutex> g(d->mu);
                    drain_pending(c, d);
                    d->have_pending_result = false;
                    rcs[(size_t)i] = probe_locked(c, d, &res[(size_t)i]);
                } catch (...) {                   // an exception leaving a thread would terminate the host process
                    rcs[(size_t)i] = CRO_ERR_INTERNAL;
                }
            });
        for (auto& t : th) t.join();
    }
    int worst = CRO_OK;
    for (int i = 0; i < n; ++i)
        if (rcs[(size_t)i] != CRO_OK && rcs[(size_t)i] != CRO_ERR_CHECKSUM) return rcs[(size_t)i];
        else if (rcs[(size_t)i] != CRO_OK) worst = rcs[(size_t)i];

    // phase 2: NVLink P2P, 1-factorised so each GPU is in exactly one pair per round
    if (n > 1 && !(o.flags & CRO_F_SKIP_P2P)) {
        int rc = enable_peers(c);
        if (rc) return rc;
        std::vector<std::unique_lock<std::mutex>> locks;
        for (int i = 0; i < n; ++i) locks.emplace_back(c->devs[(size_t)i]->mu);
        // expected checksum of each owner's first p2p_bytes
        std::vector<SweepOut> prefix((size_t)n);
        for (int i = 0; i < n; ++i) {
            Device* d = c->devs[(size_t)i].get();
            CU_TRY(c, cudaSetDevice(d->ordinal));
            if ((rc = ensure_chase(c, d))) return rc;
            CU_TRY(c, launch_expected(d->plan, std::min<uint64_t>(o.p2p_bytes, d->sweep_bytes), d->seed, d->scratch, &d->d_out[kMaxSweeps - 2], d->stream));
            c->launches++;
            CU_TRY(c, cudaMemcpyAsync(&d->h_out[kMaxSweeps - 2], &d->d_out[kMaxSweeps - 2], sizeof(SweepOut),
                                      cudaMemcpyDeviceToHost, d->stream));
            CU_TRY(c, cudaStreamSynchronize(d->stream));
            prefix[(size_t)i] = d->h_out[kMaxSweeps - 2];
        }
        for (int a = 0; a < n; ++a)
            for (int b = 0; b < n; ++b) {
                if (a == b || b >= 8 || a >= 8) continue;
                int can = 0;
                CU_TRY(c, cudaDeviceCanAccessPeer(&can, c->devs[(size_t)a]->ordinal, c->devs[(size_t)b]->ordinal));
                res[(size_t)a].p2p_access[b] = (uint8_t)can;
            }
        for (const auto& round : one_factorisation(n)) {
            // bandwidth: both directions of every pair in flight at once
            std::vector<std::pair<int, int>> directed;
            for (const auto& p : round) {
                directed.push_back({p.first, p.second});
                // CRO_P2P_UNIDIR=1 (measurement only, tools/p2p_variants.py): one direction per pair, to see what the
                // link gives when its other half is idle; the reverse direction's result slots stay zero
                if (!env_u32("CRO_P2P_UNIDIR", 0)) directed.push_back({p.second, p.first});
            }
            for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms the mappings, rep 1 is timed
                for (const auto& pr : directed) {
                    Device* a = c->devs[(size_t)pr.first].get();
                    Device* b = c->devs[(size_t)pr.second].get();
                    if (pr.first >= 8 || pr.second >= 8 || !res[(size_t)pr.first].p2p_access[pr.second]) continue;
                    CU_TRY(c, cudaSetDevice(a->ordinal));
                    CU_TRY(c, cudaEventRecord(a->ev0, a->stream));
                    // TMA bulk copies straight out of the peer's HBM (cp.async.bulk on the peer-mapped
                    // address) into this GPU's shared memory, checksummed as they land: 669 GB/s per
                    // direction with all pairs running both ways, vs 632 GB/s for LDG.128/256
                    CU_TRY(c, launch_read(a->plan, (unsigned)env_u32("CRO_P2P_READ_VARIANT", READ_TMA), b->region,
                                          std::min<uint64_t>(o.p2p_bytes, b->sweep_bytes), a->scratch, &a->d_out[0], a->stream));
                    CU_TRY(c, cudaEventRecord(a->ev1, a->stream));
                    c->launches++;
                    CU_TRY(c, cudaMemcpyAsync(&a->h_out[0], &a->d_out[0], sizeof(SweepOut), cudaMemcpyDeviceToHost, a->stream));
                }
                for (const auto& pr : directed) {
                    Device* a = c->devs[(size_t)pr.first].get();
                    if (pr.first >= 8 || pr.second >= 8 || !res[(size_t)pr.first].p2p_access[pr.second]) continue;
                    CU_TRY(c, cudaSetDevice(a->ordinal));
                    CU_TRY(c, cudaStreamSynchronize(a->stream));
                    if (rep == 0) continue;
                    float ms = 0;
                    CU_TRY(c, cudaEventElapsedTime(&ms, a->ev0, a->ev1));
                    res[(size_t)pr.first].p2p_read_ns[pr.second] = ms_to_ns(ms);
                    res[(size_t)pr.first].p2p_checksum_xor[pr.second] = a->h_out[0].x;
                    if (a->h_out[0].x != prefix[(size_t)pr.second].x || a->h_out[0].s != prefix[(size_t)pr.second].s) {
                        res[(size_t)pr.first].status = CRO_ERR_CHECKSUM;
                        worst = CRO_ERR_CHECKSUM;
                        c->set_error(std::string("NVLink read of ") + c->devs[(size_t)pr.second]->info.gpu_uuid +
                                     " from " + a->info.gpu_uuid + " does not reproduce the pattern checksum");
                    }
                }
            }
            // push leg: posted NVLink writes.  a streams its own pattern prefix through shared memory
            // (TMA bulk load from local HBM, TMA bulk store to the peer-mapped address) into the
            // SCRATCH half of b's region; b then re-reads that half locally and must find a's checksum.
            // The scratch half is rewritten by every probe's copy sweeps, so nothing needs restoring.
            if (!(o.flags & CRO_F_SKIP_P2P_WRITE)) {
                auto push_bytes = [&](const Device* a, const Device* b) {
                    return std::min<uint64_t>(std::min<uint64_t>(o.p2p_bytes, a->sweep_bytes), b->sweep_bytes);
                };
                for (int rep = 0; rep < 2; ++rep) {   // rep 0 maps the peer pages (1/16 of the bytes), rep 1 is timed
                    for (const auto& pr : directed) {
                        Device* a = c->devs[(size_t)pr.first].get();
                        Device* b = c->devs[(size_t)pr.second].get();
                        if (pr.first >= 8 || pr.second >= 8 || !res[(size_t)pr.first].p2p_access[pr.second]) continue;
                        uint64_t nb = push_bytes(a, b);
                        if (rep == 0) nb = std::max<uint64_t>(nb / 16, std::min<uint64_t>(nb, 1u << 20)) & ~uint64_t(15);
                        CU_TRY(c, cudaSetDevice(a->ordinal));
                        CU_TRY(c, cudaEventRecord(a->ev0, a->stream));
                        CU_TRY(c, launch_copy(a->plan, (unsigned)env_u32("CRO_P2P_WRITE_VARIANT", COPY_TMA),
                                              b->region + b->sweep_bytes, a->region, nb, a->scratch, a->stream));
                        CU_TRY(c, cudaEventRecord(a->ev1, a->stream));
                        c->launches++;
                    }
                    for (const auto& pr : directed) {
                        Device* a = c->devs[(size_t)pr.first].get();
                        if (pr.first >= 8 || pr.second >= 8 || !res[(size_t)pr.first].p2p_access[pr.second]) continue;
                        CU_TRY(c, cudaSetDevice(a->ordinal));
                        CU_TRY(c, cudaStreamSynchronize(a->stream));
                        if (rep == 0) continue;
                        float ms = 0;
                        CU_TRY(c, cudaEventElapsedTime(&ms, a->ev0, a->ev1));
                        res[(size_t)pr.first].p2p_write_ns[pr.second] = ms_to_ns(ms);
                    }
                }
                // every pusher has drained (stream syncs above): the receivers check what landed
                for (const auto& pr : directed) {
                    Device* a = c->devs[(size_t)pr.first].get();
                    Device* b = c->devs[(size_t)pr.second].get();
                    if (pr.first >= 8 || pr.second >= 8 || !res[(size_t)pr.first].p2p_access[pr.second]) continue;
                    CU_TRY(c, cudaSetDevice(b->ordinal));
                    CU_TRY(c, launch_read(b->plan, resolve_read_variant(CRO_READ_AUTO, push_bytes(a, b)), b->region + b->sweep_bytes,
                                          push_bytes(a, b), b->scratch, &b->d_out[1], b->stream));
                    c->launches++;
                    CU_TRY(c, cudaMemcpyAsync(&b->h_out[1], &b->d_out[1], sizeof(SweepOut), cudaMemcpyDeviceToHost, b->stream));
                }
                for (const auto& pr : directed) {
                    Device* a = c->devs[(size_t)pr.first].get();
                    Device* b = c->devs[(size_t)pr.second].get();
                    if (pr.first >= 8 || pr.second >= 8 || !res[(size_t)pr.first].p2p_access[pr.second]) continue;
                    CU_TRY(c, cudaSetDevice(b->ordinal));
                    CU_TRY(c, cudaStreamSynchronize(b->stream));
                    // prefix[] holds the checksum of min(p2p_bytes, owner's S) words; a smaller receiver
                    // region changes the byte count, in which case only the timing is reported
                    if (push_bytes(a, b) != std::min<uint64_t>(o.p2p_bytes, a->sweep_bytes)) continue;
                    if (b->h_out[1].x != prefix[(size_t)pr.first].x || b->h_out[1].s != prefix[(size_t)pr.first].s) {
                        res[(size_t)pr.first].status = CRO_ERR_CHECKSUM;
                        worst = CRO_ERR_CHECKSUM;
                        c->set_error(std::string("NVLink push from ") + a->info.gpu_uuid + " into " + b->info.gpu_uuid +
                                     " did not land the pattern checksum");
                    }
                }
            }
            // latency: dependent loads into the peer's permutation
            for (const auto& pr : directed) {
                Device* a = c->devs[(size_t)pr.first].get();
                Device* b = c->devs[(size_t)pr.second].get();
                if (pr.first >= 8 || pr.second >= 8 || !res[(size_t)pr.first].p2p_access[pr.second]) continue;
                CU_TRY(c, cudaSetDevice(a->ordinal));
                CU_TRY(c, launch_chase(b->d_chase_next, 0, o.latency_hops, a->d_chase_out, a->stream));
                c->launches++;
            }
            for (const auto& pr : directed) {
                Device* a = c->devs[(size_t)pr.first].get();
                if (pr.first >= 8 || pr.second >= 8 || !res[(size_t)pr.first].p2p_access[pr.second]) continue;
                CU_TRY(c, cudaSetDevice(a->ordinal));
                unsigned long long h[2] = {0, 0};
                CU_TRY(c, cudaMemcpyAsync(h, a->d_chase_out, sizeof h, cudaMemcpyDeviceToHost, a->stream));
                CU_TRY(c, cudaStreamSynchronize(a->stream));
                res[(size_t)pr.first].p2p_latency_ns_x16[pr.second] =
                    (uint32_t)std::min<unsigned long long>(0xFFFFFFFFull, h[1] * 16ull / std::max(1u, o.latency_hops));
            }
        }
    }

    // phase 3: ONE all-gather of the 512-byte structs over NVLink
    for (int i = 0; i < n; ++i) {
        Device* d = c->devs[(size_t)i].get();
        std::lock_guard<std::mutex> g(d->mu);
        refresh_ecc(c, d);                        // the full-box probe is rare enough to afford a fresh read
        res[(size_t)i].ecc_errors = d->ecc_uncorrected;
        int rc = publish_result(c, d, &res[(size_t)i]);
        if (rc) return rc;
    }
    const bool use_nccl = n > 1 && !(o.flags & CRO_F_SKIP_NCCL);
    if (use_nccl) {
        Nccl nc;
        int rc = load_nccl(c, &nc);
        if (rc) return rc;
        if (!c->nccl_ready) {
            std::vector<int> ords;
            for (auto& d : c->devs) ords.push_back(d->ordinal);
            c->nccl_comms.assign((size_t)n, nullptr);
            int r = nc.CommInitAll(c->nccl_comms.data(), n, ords.data());
            if (r != 0) {
                c->set_error(std::string("ncclCommInitAll: ") + (nc.GetErrorString ? nc.GetErrorString(r) : "error"));
                return CRO_ERR_NCCL;
            }
            c->nccl_ready = true;
        }
        int r = nc.GroupStart();
        for (int i = 0; r == 0 && i < n; ++i) {
            Device* d = c->devs[(size_t)i].get();
            r = nc.AllGather(d->d_result, d->d_gather, sizeof(cro_probe_result), /*ncclUint8*/ 1,
                             c->nccl_comms[(size_t)i], d->stream);
        }
        int r2 = nc.GroupEnd();
        if (r != 0 || r2 != 0) {
            c->set_error(std::string("ncclAllGather: ") + (nc.GetErrorString ? nc.GetErrorString(r ? r : r2) : "error"));
            return CRO_ERR_NCCL;
        }
        std::vector<cro_probe_result> got((size_t)n), ref((size_t)n);
        for (int i = 0; i < n; ++i) {
            Device* d = c->devs[(size_t)i].get();
            CU_TRY(c, cudaSetDevice(d->ordinal));
            CU_TRY(c, cudaMemcpyAsync(got.data(), d->d_gather, sizeof(cro_probe_result) * (size_t)n,
                                      cudaMemcpyDeviceToHost, d->stream));
            CU_TRY(c, cudaStreamSynchronize(d->stream));
            if (i == 0) ref = got;
            else if (memcmp(ref.data(), got.data(), sizeof(cro_probe_result) * (size_t)n) != 0) {
                c->set_error("all-gather result differs between rank 0 and rank " + std::to_string(i));
                return CRO_ERR_NCCL;
            }
        }
        memcpy(out, ref.data(), sizeof(cro_probe_result) * (size_t)n);
    } else {
        memcpy(out, res.data(), sizeof(cro_probe_result) * (size_t)n);
    }
    return worst;
}

}  // namespace cro
