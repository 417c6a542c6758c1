// kernels.cuh — launch wrappers for the sm_100a probe kernels.
//
// No reference counterpart: the reference's post-attach check is a UUID string
// match (internal/utils/gpus.go:54-86); these kernels are the strong check that
// takes its slot.  All arithmetic is 64-bit integer; results
// are bit-exact against the CPU restatement the tests hold (see tests/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cro {

// Per-sweep result slot in device memory, written by the last CTA to finish.
struct SweepOut {
    unsigned long long x;      // XOR of all 64-bit words
    unsigned long long s;      // wrapping sum of all 64-bit words
    unsigned long long t0;     // min %globaltimer at CTA start (ns)
    unsigned long long t1;     // max %globaltimer at CTA end   (ns)
};

// Scratch a device needs for the reductions (allocated once per device).
struct SweepScratch {
    ulonglong2*  partials;     // one (xor,sum) per CTA, >= max grid
    unsigned int* counter;     // self-resetting "CTAs done" counter
    unsigned long long* tmin;  // per-sweep timers, reset by the last CTA
    unsigned long long* tmax;
    unsigned long long* tile_ctr;  // dynamic tile counter of the TMA kernels (zeroed per launch)
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

enum : unsigned { READ_LDG = 1, READ_TMA = 2, READ_LDG256 = 3, COPY_LDG = 1, COPY_TMA = 2 };

// One-time per-device setup (smem carve-outs, occupancy → persistent grid size).
struct KernelPlan {
    LaunchCfg fill, read_ldg, read_ldg256, read_tma, copy_ldg, copy_tma, expect;
    int sm_count;
};
cudaError_t plan_kernels(int device, KernelPlan* plan);

cudaError_t launch_fill(const KernelPlan&, void* base, uint64_t bytes, uint64_t seed, cudaStream_t);
cudaError_t launch_read(const KernelPlan&, unsigned variant, const void* base, uint64_t bytes,
                        const SweepScratch&, SweepOut* out, cudaStream_t);
cudaError_t launch_copy(const KernelPlan&, unsigned variant, void* dst, const void* src,
                        uint64_t bytes, const SweepScratch&, cudaStream_t);
cudaError_t launch_expected(const KernelPlan&, uint64_t bytes, uint64_t seed,
                            const SweepScratch&, SweepOut* out, cudaStream_t);
cudaError_t launch_xor_word(void* base, uint64_t word_index, uint64_t mask, cudaStream_t);
// Pointer chase over `next` (one 8-byte slot per 128-byte line): hops loads
// with ld.relaxed.sys; out[0] = final index, out[1] = elapsed %globaltimer ns.
cudaError_t launch_chase(const unsigned long long* next, uint32_t start, uint32_t hops,
                         unsigned long long* out, cudaStream_t);

}  // namespace cro
