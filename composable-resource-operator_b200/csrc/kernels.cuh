// kernels.cuh — launch wrappers for the sm_100a probe kernels.
//
// No reference counterpart: the reference's post-attach check is a UUID string
// match (internal/utils/gpus.go:54-86); these kernels are the strong check that
// takes its slot.  All arithmetic is 64-bit integer; results
// are bit-exact against the CPU restatement the tests hold (see tests/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/croprobe.h"

namespace cro {

// Per-sweep result slot in device memory, written by the last CTA to finish.
// The checksum of a sweep over words w[0..n) is the triple
//   x = XOR of all words,  s = wrapping sum,  w = wrapping sum of w[i] * (2i + 1)
// (the third component makes the position of every word matter: swapped or
// misplaced tiles change it, which XOR and sum alone cannot see).
struct SweepOut {
    unsigned long long x;
    unsigned long long s;
    unsigned long long w;
    unsigned long long t0;      // min %globaltimer at CTA start (ns)
    unsigned long long t1;      // max %globaltimer at CTA end   (ns)
    unsigned long long stamp;   // nonce of the probe whose kernel wrote the slot (a stale slot is a failure)
    unsigned long long n_words; // words the sweep covered
    unsigned long long pad;
};
static_assert(sizeof(SweepOut) == 64, "SweepOut is one 64-byte slot");

// What a probe's kernels read from device memory instead of taking as launch
// parameters, so that ONE captured CUDA graph serves every probe: the host
// refreshes these 16 bytes (a memcpy node at the head of the graph).
struct ProbeParams {
    unsigned long long seed;    // effective pattern seed of this probe (device seed + nonce * kNonceStride)
    unsigned long long nonce;   // probe number on this device, starts at 0
};
constexpr unsigned long long kNonceStride = 0xD1B54A32D192ED03ull;   // odd: distinct nonces give distinct seeds

// Scratch a kernel needs for its reduction.  Kernels that may run CONCURRENTLY
// on one device (main stream vs the closed-form generator on the side stream)
// each own one.
struct SweepScratch {
    ulonglong4*  partials;     // one (xor, sum, wsum, -) per CTA, >= max grid
    unsigned int* counter;     // self-resetting "CTAs done" ticket
    unsigned long long* tmin;  // per-sweep timers, reset by the last CTA
    unsigned long long* tmax;
    unsigned long long* tile_ctr;  // dynamic tile counter of the TMA kernels, reset by the last CTA
};

struct LaunchCfg {
    int grid;
    int block;
    size_t smem;
};

// Pattern word i of a region: splitmix64 step of state (seed + i).
__host__ __device__ __forceinline__ unsigned long long pattern_word(unsigned long long seed,
                                                                     unsigned long long i) {
    unsigned long long z = seed + i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

enum : unsigned { READ_LDG = 1, READ_TMA = 2, READ_LDG256 = 3, COPY_LDG = 1, COPY_TMA = 2, COPY_TMA_FUSED = 3 };

// One-time per-device setup (smem carve-outs, occupancy → persistent grid size).
struct KernelPlan {
    LaunchCfg fill, read_ldg, read_ldg256, read_tma, copy_ldg, copy_tma, copy_fused, expect;
    int sm_count;
    // tuning knobs, read from the environment ONCE per device (validated; see env.hpp)
    unsigned read_tile, read_stages, read_chunk, read_dyn;
    unsigned copy_tile, copy_stages, copy_chunk, copy_dyn;
    unsigned fused_tile, fused_stages, fused_chunk, fused_threads;
};
cudaError_t plan_kernels(int device, KernelPlan* plan);

// Seed and stamp of a launch: `pp` (device pointer) when non-null — the captured graph's kernels read the 16 bytes
// the host refreshed — else the immediate `imm`.
struct Params {
    ProbeParams imm;
    const ProbeParams* pp;
};
cudaError_t launch_fill(const KernelPlan&, void* base, uint64_t bytes, const Params&,
                        const SweepScratch&, SweepOut* out, cudaStream_t);
cudaError_t launch_read(const KernelPlan&, unsigned variant, const void* base, uint64_t bytes,
                        const Params&, const SweepScratch&, SweepOut* out, cudaStream_t);
// COPY_TMA_FUSED folds every tile it moves (checksum of the SOURCE stream as read) into *out;
// the other variants leave *out alone (out may be null for them).
cudaError_t launch_copy(const KernelPlan&, unsigned variant, void* dst, const void* src, uint64_t bytes,
                        const Params&, const SweepScratch&, SweepOut* out, cudaStream_t);
cudaError_t launch_expected(const KernelPlan&, uint64_t bytes, const Params&,
                            const SweepScratch&, SweepOut* out, cudaStream_t);
cudaError_t launch_xor_word(void* base, uint64_t word_index, uint64_t mask, cudaStream_t);

// Pointer chase for NVLink latency: warp j of the one CTA follows `hops` dependent ld.relaxed.sys loads through
// table[j] (one 8-byte slot per 128-byte line, peer-resident); out[2j] = final index, out[2j+1] = %globaltimer ns.
struct ChaseArgs {
    const unsigned long long* table[CRO_MAX_DEVICES];
    unsigned start[CRO_MAX_DEVICES];
    unsigned n;           // tables to chase concurrently (one warp each)
    unsigned hops;
};
cudaError_t launch_chase(const ChaseArgs& a, unsigned long long* out, cudaStream_t);

// Slot map of one device's SweepOut array (d_out).
constexpr int kSlotFill = 0;
constexpr int kSlotSweep0 = 1;              // copies first, then reads: 1 .. 1 + C + R
constexpr int kMaxSweepsEach = 30;
constexpr int kSlotExpect = 62;             // closed form of the whole region
constexpr int kSlotPrefix = 63;             // closed form of the first p2p_bytes (what peers must read)
constexpr int kSlotP2P0 = 64;               // per peer j: 64 + 3j + {0 read, 1 push, 2 receiver re-read}
constexpr int kSlotScratch = 64 + 3 * CRO_MAX_DEVICES;   // single-sweep entry points
constexpr int kSlotCount = kSlotScratch + 4;

// The probe's verdict, computed on the device: fills *out (the all-gather send buffer) from the template
// (identity, staged at init), the sweep slots and the closed form.  One CTA.
struct FinalizeArgs {
    const cro_probe_result* tmpl;
    cro_probe_result* out;
    const SweepOut* slots;
    const ProbeParams* pp;
    unsigned long long sweep_bytes;
    unsigned read_sweeps, copy_sweeps;
    unsigned read_variant, copy_variant;
    unsigned fused;          // copy sweeps carry a checksum of their source stream
};
cudaError_t launch_finalize(const FinalizeArgs& a, cudaStream_t);

// NVLink part of the verdict: folds the per-peer slots of this device (and the peers' slots it must agree with)
// into out->p2p_* and out->status.  One CTA.
struct P2PFinalizeArgs {
    cro_probe_result* out;
    const SweepOut* slots;                        // this device's
    const SweepOut* peer_slots[CRO_MAX_DEVICES];  // peer j's slot array (peer-mapped), null = no such peer
    const unsigned long long* chase_out;          // [2j] end index, [2j+1] ns, indexed by peer
    unsigned chase_expect[CRO_MAX_DEVICES];       // where the chase into peer j must end
    unsigned n;                                   // devices managed
    unsigned self;
    unsigned hops;
    unsigned have_push;
    unsigned push_folded;                         // the push kernel carries a checksum of its source (COPY_TMA_FUSED)
    unsigned long long p2p_bytes;
    unsigned long long stamp;                     // nonce every p2p slot this device wrote must carry
    unsigned long long peer_stamp[CRO_MAX_DEVICES];   // ... and the nonce of peer j's closed-form slot
};
cudaError_t launch_p2p_finalize(const P2PFinalizeArgs& a, cudaStream_t);

}  // namespace cro
