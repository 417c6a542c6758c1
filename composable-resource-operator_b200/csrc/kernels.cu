// kernels.cu — hand-written sm_100a kernels of the post-attach HBM / NVLink probe.
//
// The reference has no device code at all (its check is the UUID membership
// test at internal/utils/gpus.go:54-86); these kernels are new work in that
// slot.  They are HBM-bound integer sweeps:
//
//   hbm_fill          S bytes written   w[i] = splitmix64-step(seed + i)
//   hbm_read_*        S bytes read      (XOR-fold, wrapping sum) of all words
//   hbm_copy_*        2S bytes moved
//   hbm_expected      0 bytes           the same checksum from the closed form
//   chase             pointer chase over a peer-resident permutation (latency)
//
// Two data paths per sweep: 128-bit ld.global.nc / st.global vector accesses
// (also used on peer-mapped pointers for the NVLink probe), and 1-D TMA bulk
// copies (cp.async.bulk + mbarrier) through a shared-memory ring.
// No tensor cores: there is no contraction anywhere on this path.
#include "kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace cro {

// ---------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Batched streaming loads: ONE asm block so ptxas cannot interleave the
// consumers between the loads — every thread keeps the whole batch (128 B) in
// flight.  STRIDE is the byte distance between a thread's consecutive vectors.
template <int STRIDE>
__device__ __forceinline__ void ldg128_x8(const void* p, unsigned long long (&a)[8],
                                          unsigned long long (&b)[8]) {
    asm volatile(
        "ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%16];\n"
        "ld.global.nc.L1::no_allocate.v2.u64 {%2,%3}, [%16+%17];\n"
        "ld.global.nc.L1::no_allocate.v2.u64 {%4,%5}, [%16+%18];\n"
        "ld.global.nc.L1::no_allocate.v2.u64 {%6,%7}, [%16+%19];\n"
        "ld.global.nc.L1::no_allocate.v2.u64 {%8,%9}, [%16+%20];\n"
        "ld.global.nc.L1::no_allocate.v2.u64 {%10,%11}, [%16+%21];\n"
        "ld.global.nc.L1::no_allocate.v2.u64 {%12,%13}, [%16+%22];\n"
        "ld.global.nc.L1::no_allocate.v2.u64 {%14,%15}, [%16+%23];\n"
        : "=l"(a[0]), "=l"(b[0]), "=l"(a[1]), "=l"(b[1]), "=l"(a[2]), "=l"(b[2]), "=l"(a[3]),
          "=l"(b[3]), "=l"(a[4]), "=l"(b[4]), "=l"(a[5]), "=l"(b[5]), "=l"(a[6]), "=l"(b[6]),
          "=l"(a[7]), "=l"(b[7])
        : "l"(p), "n"(STRIDE), "n"(2 * STRIDE), "n"(3 * STRIDE), "n"(4 * STRIDE), "n"(5 * STRIDE),
          "n"(6 * STRIDE), "n"(7 * STRIDE));
}
// 256-bit flavour (sm_100+: LDG.E.256): four 32-byte vectors per thread.
template <int STRIDE>
__device__ __forceinline__ void ldg256_x4(const void* p, unsigned long long (&w)[16]) {
    asm volatile(
        "ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%16];\n"
        "ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%4,%5,%6,%7}, [%16+%17];\n"
        "ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%8,%9,%10,%11}, [%16+%18];\n"
        "ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%12,%13,%14,%15}, [%16+%19];\n"
        : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]), "=l"(w[4]), "=l"(w[5]), "=l"(w[6]),
          "=l"(w[7]), "=l"(w[8]), "=l"(w[9]), "=l"(w[10]), "=l"(w[11]), "=l"(w[12]), "=l"(w[13]),
          "=l"(w[14]), "=l"(w[15])
        : "l"(p), "n"(STRIDE), "n"(2 * STRIDE), "n"(3 * STRIDE));
}

__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

__device__ __forceinline__ unsigned long long lo64(const uint4& v) {
    return (unsigned long long)v.x | ((unsigned long long)v.y << 32);
}
__device__ __forceinline__ unsigned long long hi64(const uint4& v) {
    return (unsigned long long)v.z | ((unsigned long long)v.w << 32);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 1-D TMA: shared -> global, tracked by bulk async-groups.
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
// L2 eviction-priority policies for the bulk copies (streaming data is touched once).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gsrc, uint32_t bytes,
                                                 uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_1d_hint(void* gdst, const void* smem_src, uint32_t bytes,
                                                  uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
// CTA reduction + "last CTA publishes" epilogue shared by read / expected.
// XOR and wrapping add are associative and commutative, so the result is
// independent of grid shape and scheduling: bit-exact by construction.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void publish(unsigned long long x, unsigned long long s,
                                        unsigned long long t_start, const SweepScratch sc,
                                        SweepOut* out) {
    __shared__ unsigned long long sx[32], ss[32];
    __shared__ bool is_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (blockDim.x + 31) >> 5;
    if (lane == 0) { sx[warp] = x; ss[warp] = s; }
    __syncthreads();
    if (warp == 0) {
        x = lane < nwarps ? sx[lane] : 0ull;
        s = lane < nwarps ? ss[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            x ^= __shfl_xor_sync(0xffffffffu, x, o);
            s += __shfl_xor_sync(0xffffffffu, s, o);
        }
        if (lane == 0) {
            sc.partials[blockIdx.x] = make_ulonglong2(x, s);
            atomicMin(sc.tmin, t_start);
            atomicMax(sc.tmax, globaltimer_ns());
            __threadfence();
            unsigned ticket = atomicAdd(sc.counter, 1u);
            is_last = (ticket == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    x = 0; s = 0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
        ulonglong2 p = __ldcg(&sc.partials[i]);
        x ^= p.x; s += p.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    __syncthreads();
    if (lane == 0) { sx[warp] = x; ss[warp] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        x = 0; s = 0;
        for (int w = 0; w < nwarps; ++w) { x ^= sx[w]; s += ss[w]; }
        out->x = x;
        out->s = s;
        out->t0 = *((volatile unsigned long long*)sc.tmin);
        out->t1 = *((volatile unsigned long long*)sc.tmax);
        *sc.counter = 0u;
        *sc.tmin = ~0ull;
        *sc.tmax = 0ull;
        __threadfence();
    }
}

// ---------------------------------------------------------------------------
// hbm_fill: S bytes written.  Thread t of a tile stores vectors t, t+T, ...
// so each warp-level store instruction covers 512 contiguous bytes.
// ---------------------------------------------------------------------------
template <int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
hbm_fill_kernel(uint4* __restrict__ base, unsigned long long n_vec, unsigned long long seed) {
    const unsigned long long tile_vecs = (unsigned long long)THREADS * UNROLL;
    const unsigned long long n_tiles = n_vec / tile_vecs;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long v0 = tile * tile_vecs + threadIdx.x;
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const unsigned long long v = v0 + (unsigned long long)j * THREADS;
            const unsigned long long a = pattern_word(seed, 2 * v);
            const unsigned long long b = pattern_word(seed, 2 * v + 1);
            stg_stream(base + v, make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b,
                                            (unsigned)(b >> 32)));
        }
    }
    // ragged tail (S not a multiple of the tile): plain grid-stride
    for (unsigned long long v = n_tiles * tile_vecs + (unsigned long long)blockIdx.x * THREADS +
                                threadIdx.x;
         v < n_vec; v += (unsigned long long)gridDim.x * THREADS) {
        const unsigned long long a = pattern_word(seed, 2 * v);
        const unsigned long long b = pattern_word(seed, 2 * v + 1);
        stg_stream(base + v,
                   make_uint4((unsigned)a, (unsigned)(a >> 32), (unsigned)b, (unsigned)(b >> 32)));
    }
}

// ---------------------------------------------------------------------------
// hbm_read (LDG path): UNROLL independent 128-bit ld.global.nc per thread in
// flight, read-only path, no L1 allocation.  Also the NVLink P2P read kernel
// (base may be a peer-mapped pointer).
// ---------------------------------------------------------------------------
template <int THREADS, bool WIDE>
__global__ void __launch_bounds__(THREADS)
hbm_read_ldg_kernel(const uint4* __restrict__ base, unsigned long long n_vec, SweepScratch sc,
                    SweepOut* out) {
    const unsigned long long t_start = globaltimer_ns();
    unsigned long long x0 = 0, x1 = 0, s0 = 0, s1 = 0;
    constexpr unsigned long long tile_vecs = (unsigned long long)THREADS * 8;  // 128 B / thread
    const unsigned long long n_tiles = n_vec / tile_vecs;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (WIDE) {
            // thread t owns 32-byte vectors t, t+T, t+2T, t+3T of the tile
            const unsigned char* p = reinterpret_cast<const unsigned char*>(base + tile * tile_vecs) +
                                     (size_t)threadIdx.x * 32;
            unsigned long long w[16];
            ldg256_x4<THREADS * 32>(p, w);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                x0 ^= w[j]; x1 ^= w[j + 1]; s0 += w[j]; s1 += w[j + 1];
            }
        } else {
            const uint4* p = base + tile * tile_vecs + threadIdx.x;
            unsigned long long a[8], b[8];
            ldg128_x8<THREADS * 16>(p, a, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x0 ^= a[j]; x1 ^= b[j]; s0 += a[j]; s1 += b[j];
            }
        }
    }
    for (unsigned long long i = n_tiles * tile_vecs + (unsigned long long)blockIdx.x * THREADS +
                                threadIdx.x;
         i < n_vec; i += (unsigned long long)gridDim.x * THREADS) {
        const uint4 v = ldg_stream(base + i);
        const unsigned long long a = lo64(v), b = hi64(v);
        x0 ^= a; x1 ^= b; s0 += a; s1 += b;
    }
    publish(x0 ^ x1, s0 + s1, t_start, sc, out);
}

// ---------------------------------------------------------------------------
// hbm_read (TMA path): warp 0 / lane 0 is the producer, issuing 1-D bulk
// copies of `tile_bytes` into a `stages`-deep shared-memory ring; the consumer
// warps fold each landed tile with conflict-free LDS.128 and hand the slot
// back through an "empty" mbarrier.  Bytes in flight per SM = stages*tile.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
hbm_read_tma_kernel(const unsigned char* __restrict__ base, unsigned long long bytes,
                    unsigned tile_bytes, unsigned stages, unsigned chunk,
                    unsigned long long* tile_ctr, SweepScratch sc, SweepOut* out) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t full_bar[16];
    __shared__ __align__(8) uint64_t empty_bar[16];
    __shared__ unsigned long long tile_of[16];
    const unsigned long long t_start = globaltimer_ns();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned n_cons_warps = (blockDim.x >> 5) - 1;
    const unsigned long long n_tiles = (bytes + tile_bytes - 1) / tile_bytes;
    const unsigned hint = chunk >> 16;   // bit 0: L2 evict_first on the loads
    chunk &= 0xFFFFu;

    if (threadIdx.x == 0) {
        for (unsigned s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], n_cons_warps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    unsigned long long x0 = 0, x1 = 0, s0 = 0, s1 = 0;
    constexpr unsigned long long kEnd = ~0ull;
    if (warp == 0) {
        if (lane == 0) {
            // Producer.  Tiles come from a device-wide atomic counter (dynamic:
            // fast SMs take more tiles, the in-flight window stays compact) or
            // from static striding when tile_ctr is null.
            unsigned stage = 0, phase = 0;
            unsigned long long cur = 0, end = 0;   // claimed-but-unissued tiles [cur, end)
            for (unsigned long long k = 0;; ++k) {
                unsigned long long tile;
                if (tile_ctr) {
                    if (cur == end) { cur = atomicAdd(tile_ctr, (unsigned long long)chunk); end = cur + chunk; }
                    tile = cur++;
                } else {
                    tile = blockIdx.x + k * (unsigned long long)gridDim.x;
                }
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (tile >= n_tiles) {
                    tile_of[stage] = kEnd;
                    mbar_arrive(&full_bar[stage]);      // completes the phase with no bytes
                    break;
                }
                tile_of[stage] = tile;
                const unsigned long long off = tile * tile_bytes;
                const unsigned long long left = bytes - off;
                const unsigned nb = left < tile_bytes ? (unsigned)left : tile_bytes;
                mbar_expect_tx(&full_bar[stage], nb);  // release: publishes tile_of[stage]
                if (hint & 1u)
                    tma_load_1d_hint(ring + (size_t)stage * tile_bytes, base + off, nb, &full_bar[stage],
                                     l2_policy_evict_first());
                else
                    tma_load_1d(ring + (size_t)stage * tile_bytes, base + off, nb, &full_bar[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        const unsigned ctid = threadIdx.x - 32;
        const unsigned n_cons = n_cons_warps * 32;
        unsigned stage = 0, phase = 0;
        for (;;) {
            mbar_wait(&full_bar[stage], phase);
            const unsigned long long tile = *reinterpret_cast<volatile unsigned long long*>(&tile_of[stage]);
            if (tile == kEnd) break;
            const unsigned long long off = tile * tile_bytes;
            const unsigned long long left = bytes - off;
            const unsigned nvec = (left < tile_bytes ? (unsigned)left : tile_bytes) >> 4;
            const uint4* sp = reinterpret_cast<const uint4*>(ring + (size_t)stage * tile_bytes);
            unsigned i = ctid;
            // 4 independent LDS.128 per trip
            for (; i + 3 * n_cons < nvec; i += 4 * n_cons) {
                const uint4 a = sp[i], b = sp[i + n_cons], c = sp[i + 2 * n_cons],
                            d = sp[i + 3 * n_cons];
                x0 ^= lo64(a) ^ lo64(c); x1 ^= hi64(a) ^ hi64(c);
                s0 += lo64(a) + lo64(c); s1 += hi64(a) + hi64(c);
                x0 ^= lo64(b) ^ lo64(d); x1 ^= hi64(b) ^ hi64(d);
                s0 += lo64(b) + lo64(d); s1 += hi64(b) + hi64(d);
            }
            for (; i < nvec; i += n_cons) {
                const uint4 a = sp[i];
                x0 ^= lo64(a); x1 ^= hi64(a); s0 += lo64(a); s1 += hi64(a);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
    }
    publish(x0 ^ x1, s0 + s1, t_start, sc, out);
}

// ---------------------------------------------------------------------------
// hbm_copy (LDG/STG path)
// ---------------------------------------------------------------------------
template <int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
hbm_copy_ldg_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src,
                    unsigned long long n_vec) {
    const unsigned long long tile_vecs = (unsigned long long)THREADS * UNROLL;
    const unsigned long long n_tiles = n_vec / tile_vecs;
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long o = tile * tile_vecs + threadIdx.x;
        static_assert(UNROLL == 8, "batched loader is written for 8 vectors per thread");
        unsigned long long a[8], b[8];
        ldg128_x8<THREADS * 16>(src + o, a, b);
#pragma unroll
        for (int j = 0; j < UNROLL; ++j)
            stg_stream(dst + o + j * THREADS,
                       make_uint4((unsigned)a[j], (unsigned)(a[j] >> 32), (unsigned)b[j],
                                  (unsigned)(b[j] >> 32)));
    }
    for (unsigned long long i = n_tiles * tile_vecs + (unsigned long long)blockIdx.x * THREADS +
                                threadIdx.x;
         i < n_vec; i += (unsigned long long)gridDim.x * THREADS)
        stg_stream(dst + i, ldg_stream(src + i));
}

// ---------------------------------------------------------------------------
// hbm_copy (TMA path): one thread per CTA drives everything.  Tiles go
// global -> smem (bulk load, mbarrier) -> global (bulk store, bulk group);
// data never touches the register file.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32, 1)
hbm_copy_tma_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src,
                    unsigned long long bytes, unsigned tile_bytes, unsigned stages, unsigned chunk,
                    unsigned long long* tile_ctr) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t full_bar[16];
    __shared__ unsigned long long tile_of[16];
    if (threadIdx.x != 0) return;
    for (unsigned s = 0; s < stages; ++s) mbar_init(&full_bar[s], 1);
    mbar_fence_init();
    // bit 0: evict_first loads, bit 1: evict_first stores, bit 2: evict_last stores
    const unsigned hint = chunk >> 16;
    chunk &= 0xFFFFu;
    const uint64_t pol_first = l2_policy_evict_first(), pol_last = l2_policy_evict_last();

    const unsigned long long n_tiles = (bytes + tile_bytes - 1) / tile_bytes;
    unsigned long long fetched = 0;   // tiles this CTA has asked for
    bool dry = false;                 // the counter ran past the last tile
    unsigned long long cur = 0, end = 0;   // claimed-but-unissued tiles [cur, end)
    auto fetch = [&]() -> unsigned long long {
        unsigned long long t;
        if (tile_ctr) {
            if (cur == end) { cur = atomicAdd(tile_ctr, (unsigned long long)chunk); end = cur + chunk; }
            t = cur++;
        } else {
            t = blockIdx.x + fetched * (unsigned long long)gridDim.x;
        }
        ++fetched;
        if (t >= n_tiles) dry = true;
        return t;
    };
    auto tile_len = [&](unsigned long long off) {
        const unsigned long long left = bytes - off;
        return left < tile_bytes ? (unsigned)left : tile_bytes;
    };
    auto issue_load = [&](unsigned st, unsigned long long tile) {
        const unsigned long long off = tile * tile_bytes;
        const unsigned nb = tile_len(off);
        tile_of[st] = tile;
        mbar_expect_tx(&full_bar[st], nb);
        if (hint & 1u) tma_load_1d_hint(ring + (size_t)st * tile_bytes, src + off, nb, &full_bar[st], pol_first);
        else tma_load_1d(ring + (size_t)st * tile_bytes, src + off, nb, &full_bar[st]);
    };
    // prologue: fill the ring
    unsigned long long loaded = 0;
    for (unsigned s = 0; s < stages && !dry; ++s) {
        const unsigned long long t = fetch();
        if (!dry) { issue_load(s, t); ++loaded; }
    }
    for (unsigned long long k = 0; k < loaded; ++k) {
        const unsigned st = (unsigned)(k % stages);
        const unsigned phase = (unsigned)((k / stages) & 1ull);
        mbar_wait(&full_bar[st], phase);
        const unsigned long long off = tile_of[st] * tile_bytes;
        if (hint & 6u)
            tma_store_1d_hint(dst + off, ring + (size_t)st * tile_bytes, tile_len(off), (hint & 2u) ? pol_first : pol_last);
        else
            tma_store_1d(dst + off, ring + (size_t)st * tile_bytes, tile_len(off));
        tma_commit();
        // refill the slot whose store was issued one trip ago
        if (k >= 1 && !dry) {
            const unsigned long long t = fetch();
            if (!dry) {
                tma_wait_read<1>();   // all but the newest store have finished reading smem
                issue_load((unsigned)((k - 1) % stages), t);
                ++loaded;
            }
        }
    }
    tma_wait_all();
}

// ---------------------------------------------------------------------------
// hbm_expected: the checksum of the pattern from its closed form, no HBM.
// An independent generator: a fill or read bug cannot cancel out.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hbm_expected_kernel(unsigned long long n_words, unsigned long long seed, SweepScratch sc,
                    SweepOut* out) {
    const unsigned long long t_start = globaltimer_ns();
    unsigned long long x = 0, s = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
         i < n_words; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long w = pattern_word(seed, i);
        x ^= w; s += w;
    }
    publish(x, s, t_start, sc, out);
}

__global__ void xor_word_kernel(unsigned long long* base, unsigned long long idx,
                                unsigned long long mask) {
    base[idx] ^= mask;
}

// Latency: dependent loads, system scope, one slot per 128-byte line.
__global__ void chase_kernel(const unsigned long long* __restrict__ next, unsigned start,
                             unsigned hops, unsigned long long* out) {
    unsigned long long idx = start;
    // warm the TLB / first line
    {
        unsigned long long v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(next + idx * 16));
        if (v == ~0ull) out[2] = v;   // keep the warm-up load alive
    }
    const unsigned long long t0 = globaltimer_ns();
    for (unsigned h = 0; h < hops; ++h) {
        unsigned long long v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(next + idx * 16));
        idx = v;
    }
    const unsigned long long t1 = globaltimer_ns();
    out[0] = idx;
    out[1] = t1 - t0;
}

// ---------------------------------------------------------------------------
// host side: plan + launch wrappers
// ---------------------------------------------------------------------------
namespace {
constexpr int kFillThreads = 512, kFillUnroll = 4;
constexpr int kReadThreads = 512;
constexpr int kCopyThreads = 512, kCopyUnroll = 8;

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}

struct TmaTune { unsigned tile, stages, threads, chunk; };
TmaTune read_tma_tune() {
    TmaTune t;
    // defaults = best point of the sweeps in profiles/r01_sweeps.md
    t.tile = (unsigned)env_int("CRO_TMA_READ_TILE", 32768);
    t.stages = (unsigned)env_int("CRO_TMA_READ_STAGES", 4);
    t.threads = (unsigned)env_int("CRO_TMA_READ_THREADS", 160);  // 1 producer + 4 consumer warps
    t.chunk = (unsigned)env_int("CRO_TMA_READ_CHUNK", 1);
    if (t.chunk < 1) t.chunk = 1;
    t.chunk = (t.chunk & 0xFFFFu) | ((unsigned)env_int("CRO_TMA_READ_HINT", 0) << 16);
    if (t.stages > 16) t.stages = 16;
    return t;
}
TmaTune copy_tma_tune() {
    TmaTune t;
    t.tile = (unsigned)env_int("CRO_TMA_COPY_TILE", 32768);
    t.stages = (unsigned)env_int("CRO_TMA_COPY_STAGES", 4);
    t.threads = 32;
    t.chunk = (unsigned)env_int("CRO_TMA_COPY_CHUNK", 1);
    if (t.chunk < 1) t.chunk = 1;
    t.chunk = (t.chunk & 0xFFFFu) | ((unsigned)env_int("CRO_TMA_COPY_HINT", 0) << 16);
    if (t.stages > 16) t.stages = 16;
    return t;
}
}  // namespace

cudaError_t plan_kernels(int device, KernelPlan* plan) {
    cudaError_t e;
    int sms = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess)
        return e;
    plan->sm_count = sms;
    int occ = 0;

    auto fill = hbm_fill_kernel<kFillThreads, kFillUnroll>;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fill, kFillThreads, 0)) != cudaSuccess)
        return e;
    plan->fill = {sms * (occ > 0 ? occ : 1) * env_int("CRO_FILL_WAVES", 64), kFillThreads, 0};

    auto rd = hbm_read_ldg_kernel<kReadThreads, false>;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rd, kReadThreads, 0)) != cudaSuccess)
        return e;
    plan->read_ldg = {sms * (occ > 0 ? occ : 1) * env_int("CRO_READ_WAVES", 1), kReadThreads, 0};
    auto rdw = hbm_read_ldg_kernel<kReadThreads, true>;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rdw, kReadThreads, 0)) != cudaSuccess)
        return e;
    plan->read_ldg256 = {sms * (occ > 0 ? occ : 1) * env_int("CRO_READ_WAVES", 1), kReadThreads, 0};

    auto cp = hbm_copy_ldg_kernel<kCopyThreads, kCopyUnroll>;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cp, kCopyThreads, 0)) != cudaSuccess)
        return e;
    plan->copy_ldg = {sms * (occ > 0 ? occ : 1) * env_int("CRO_COPY_WAVES", 128), kCopyThreads, 0};

    {
        const TmaTune t = read_tma_tune();
        const size_t smem = (size_t)t.tile * t.stages;
        if ((e = cudaFuncSetAttribute(hbm_read_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem)) != cudaSuccess)
            return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hbm_read_tma_kernel,
                                                               (int)t.threads, smem)) != cudaSuccess)
            return e;
        plan->read_tma = {sms * (occ > 0 ? occ : 1) * env_int("CRO_TMA_READ_WAVES", 1), (int)t.threads, smem};
    }
    {
        const TmaTune t = copy_tma_tune();
        const size_t smem = (size_t)t.tile * t.stages;
        if ((e = cudaFuncSetAttribute(hbm_copy_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem)) != cudaSuccess)
            return e;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hbm_copy_tma_kernel, 32, smem)) !=
            cudaSuccess)
            return e;
        plan->copy_tma = {sms * (occ > 0 ? occ : 1) * env_int("CRO_TMA_COPY_WAVES", 1), 32, smem};
    }
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hbm_expected_kernel, 256, 0)) !=
        cudaSuccess)
        return e;
    plan->expect = {sms * (occ > 0 ? occ : 1), 256, 0};
    return cudaSuccess;
}

// No more CTAs than there are tiles: a small sweep should not pay for a grid sized for 4 GiB.
static int clamp_grid(int planned, uint64_t bytes, uint64_t tile_bytes) {
    const uint64_t tiles = (bytes + tile_bytes - 1) / tile_bytes;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)planned, tiles));
}

cudaError_t launch_fill(const KernelPlan& p, void* base, uint64_t bytes, uint64_t seed,
                        cudaStream_t st) {
    hbm_fill_kernel<kFillThreads, kFillUnroll><<<clamp_grid(p.fill.grid, bytes, kFillThreads * kFillUnroll * 16), p.fill.block, 0, st>>>(
        static_cast<uint4*>(base), bytes >> 4, seed);
    return cudaGetLastError();
}

cudaError_t launch_read(const KernelPlan& p, unsigned variant, const void* base, uint64_t bytes,
                        const SweepScratch& sc, SweepOut* out, cudaStream_t st) {
    if (variant == READ_TMA) {
        const TmaTune t = read_tma_tune();
        unsigned long long* ctr = nullptr;
        if (env_int("CRO_TMA_READ_DYN", 1) && sc.tile_ctr) {
            ctr = sc.tile_ctr;
            cudaError_t e = cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st);
            if (e != cudaSuccess) return e;
        }
        hbm_read_tma_kernel<<<clamp_grid(p.read_tma.grid, bytes, t.tile), p.read_tma.block, p.read_tma.smem, st>>>(
            static_cast<const unsigned char*>(base), bytes, t.tile, t.stages, t.chunk, ctr, sc, out);
    } else if (variant == READ_LDG256) {
        hbm_read_ldg_kernel<kReadThreads, true><<<clamp_grid(p.read_ldg256.grid, bytes, kReadThreads * 128), p.read_ldg256.block, 0, st>>>(
            static_cast<const uint4*>(base), bytes >> 4, sc, out);
    } else {
        hbm_read_ldg_kernel<kReadThreads, false><<<clamp_grid(p.read_ldg.grid, bytes, kReadThreads * 128), p.read_ldg.block, 0, st>>>(
            static_cast<const uint4*>(base), bytes >> 4, sc, out);
    }
    return cudaGetLastError();
}

cudaError_t launch_copy(const KernelPlan& p, unsigned variant, void* dst, const void* src,
                        uint64_t bytes, const SweepScratch& sc, cudaStream_t st) {
    if (variant == COPY_TMA) {
        const TmaTune t = copy_tma_tune();
        unsigned long long* ctr = nullptr;
        if (env_int("CRO_TMA_COPY_DYN", 1) && sc.tile_ctr) {
            ctr = sc.tile_ctr;
            cudaError_t e = cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), st);
            if (e != cudaSuccess) return e;
        }
        hbm_copy_tma_kernel<<<clamp_grid(p.copy_tma.grid, bytes, t.tile), p.copy_tma.block, p.copy_tma.smem, st>>>(
            static_cast<unsigned char*>(dst), static_cast<const unsigned char*>(src), bytes, t.tile,
            t.stages, t.chunk, ctr);
    } else {
        hbm_copy_ldg_kernel<kCopyThreads, kCopyUnroll><<<clamp_grid(p.copy_ldg.grid, bytes, kCopyThreads * kCopyUnroll * 16), p.copy_ldg.block, 0, st>>>(
            static_cast<uint4*>(dst), static_cast<const uint4*>(src), bytes >> 4);
    }
    return cudaGetLastError();
}

cudaError_t launch_expected(const KernelPlan& p, uint64_t bytes, uint64_t seed,
                            const SweepScratch& sc, SweepOut* out, cudaStream_t st) {
    hbm_expected_kernel<<<p.expect.grid, p.expect.block, 0, st>>>(bytes >> 3, seed, sc, out);
    return cudaGetLastError();
}

cudaError_t launch_xor_word(void* base, uint64_t word_index, uint64_t mask, cudaStream_t st) {
    xor_word_kernel<<<1, 1, 0, st>>>(static_cast<unsigned long long*>(base), word_index, mask);
    return cudaGetLastError();
}

cudaError_t launch_chase(const unsigned long long* next, uint32_t start, uint32_t hops,
                         unsigned long long* out, cudaStream_t st) {
    chase_kernel<<<1, 1, 0, st>>>(next, start, hops, out);
    return cudaGetLastError();
}

}  // namespace cro
