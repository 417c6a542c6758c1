/*
 * croprobe.h — C ABI of libcroprobe, the B200-native post-attach device probe
 * and spec-emit path for the composable-resource operator.
 *
 * This header is the drop-in boundary.  Everything here is
 * plain C: opaque context pointer, caller-allocated output buffers, integer
 * return codes.  No torch / C++ types cross it.  A Go host binds it with cgo
 * (see INTEGRATION.md and composable-resource-operator_b200/go/internal/cuda);
 * the tests and bench.py bind it with ctypes.
 *
 * Each entry point names the reference interface (path:line under the
 * reference tree) whose slot in the reconcile loop it fills or replaces.
 *
 * Threading: every entry point is re-entrant.  Calls that touch a device take
 * that device's mutex and call cudaSetDevice themselves, so they may be issued
 * from any OS thread (cgo migrates goroutines between threads).  No callbacks.
 * Ownership: cro_ctx is library-owned (cro_probe_destroy frees it); every
 * other buffer is caller-owned and is not retained after the call returns.
 */
#ifndef CROPROBE_H_
#define CROPROBE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRO_ABI_VERSION 2u

/* ---- return codes (0 ok, <0 error; text via cro_strerror) ---------------- */
#define CRO_OK                  0
#define CRO_ERR_INVALID_ARG    -1
#define CRO_ERR_ABI_MISMATCH   -2
#define CRO_ERR_NO_DEVICE      -3   /* zero CUDA devices: not an error for enumerate (n=0) */
#define CRO_ERR_CUDA           -4   /* a CUDA runtime call failed; see cro_last_error */
#define CRO_ERR_OOM            -5   /* sweep buffer could not be allocated (device busy) */
#define CRO_ERR_CHECKSUM       -6   /* HBM sweep checksum differs from the closed form   */
#define CRO_ERR_BUFFER_SMALL   -7   /* caller buffer too small; *len holds the need      */
#define CRO_ERR_NCCL           -8
#define CRO_ERR_DEADLINE       -9
#define CRO_ERR_UNSUPPORTED   -10   /* e.g. unsupported field in a query string          */
#define CRO_ERR_PARSE         -11   /* malformed CSV / JSON handed to a parser           */
#define CRO_ERR_EXEC          -12   /* the enumerate command reported stderr / exec error */
#define CRO_ERR_P2P           -13
#define CRO_ERR_INTERNAL      -14

/* ---- option flags -------------------------------------------------------- */
#define CRO_F_SKIP_COPY       0x0001u  /* do not run the hbm_copy sweeps           */
#define CRO_F_SKIP_P2P        0x0002u  /* cro_probe_all: no NVLink rounds          */
#define CRO_F_SKIP_NCCL       0x0004u  /* cro_probe_all: host gather, no NCCL      */
#define CRO_F_NO_NVML         0x0008u  /* identity from /proc + CUDA runtime only  */
#define CRO_F_VERIFY_COPY     0x0010u  /* accepted, no effect: since ABI 2 every copy destination is re-read and
                                          compared inside the probe (copy_verified)                            */
#define CRO_F_LAZY_ALLOC      0x0020u  /* allocate sweep buffers at first probe    */
#define CRO_F_DEGRADE_ON_OOM  0x0040u  /* busy device: halve S (>= 64 MiB) instead of failing;
                                          the result's sweep_bytes says what was swept  */
#define CRO_F_SKIP_P2P_WRITE  0x0080u  /* cro_probe_all: peer reads only, no push leg */
#define CRO_F_TEST_INJECT     0x0100u  /* fault injection INSIDE the probe (tests of the device-side verdict): after sweep
                                          number test_inject_after (launch order: 0 = the fill, 1.. = copies, then reads)
                                          XOR test_inject_mask into 64-bit word test_inject_word of the region (half A is
                                          words [0, S/8), half B [S/8, 2S/8)) */

/* read-sweep kernel variants */
#define CRO_READ_AUTO   0u   /* by sweep size: 256-bit LDG up to 512 MiB, the TMA ring above (measured crossover) */
#define CRO_READ_LDG    1u   /* ld.global.nc.L1::no_allocate 128-bit, unrolled      */
#define CRO_READ_TMA    2u   /* cp.async.bulk (1-D TMA) smem ring + LDS.128 reduce  */
#define CRO_COPY_AUTO   0u
#define CRO_COPY_LDG    1u
#define CRO_COPY_TMA    2u   /* bulk load -> smem -> bulk store, no register pass   */
#define CRO_COPY_TMA_FUSED 3u /* the same, and consumer warps fold every tile out of shared memory: the sweep
                                yields the checksum of its source as read (the probe's default)          */
#define CRO_READ_LDG256 3u   /* 256-bit LDG flavour                                  */

#define CRO_MAX_DEVICES 16

typedef struct cro_ctx cro_ctx;

/* Options for cro_probe_init.  Zero means "default" for every field. */
typedef struct cro_opts {
    uint32_t abi_version;          /* must be CRO_ABI_VERSION                               */
    uint32_t flags;                /* CRO_F_*                                               */
    uint64_t sweep_bytes;          /* S, per-device sweep region; default 4 GiB             */
    uint64_t p2p_bytes;            /* S_p2p per directed pair; default 1 GiB (<= S)         */
    uint64_t seed_base;            /* default 0x00C0FFEE00000000; seed = base | minor       */
    uint32_t read_sweeps;          /* default 5                                             */
    uint32_t copy_sweeps;          /* default 5                                             */
    uint32_t latency_hops;         /* pointer-chase hops per directed pair; default 1024    */
    uint32_t read_variant;         /* CRO_READ_*                                            */
    uint32_t copy_variant;         /* CRO_COPY_*                                            */
    int32_t  deadline_ms;          /* per-call deadline, 0 = none (Go ctx cannot cross cgo) */
    int32_t  n_devices;            /* 0 = every visible CUDA device                         */
    int32_t  devices[CRO_MAX_DEVICES]; /* CUDA ordinals to manage                           */
    uint32_t rank_base;            /* one-process-per-GPU hosts: this process's first rank  */
    uint32_t world_override;       /* ... and the job's world size (0 = devices managed)    */
    uint32_t test_inject_after;    /* CRO_F_TEST_INJECT: see above                          */
    uint32_t reserved0;
    uint64_t test_inject_word;
    uint64_t test_inject_mask;
} cro_opts;

/*
 * Identity of one device in the spellings the reference consumes
 * (internal/utils/gpus.go:878-919 parses `nvidia-smi --query-gpu=...` CSV;
 * :1014-1089 parses /proc/driver/nvidia/gpus/<bus>/information).
 */
typedef struct cro_dev_info {
    int32_t  cuda_ordinal;
    int32_t  device_minor;         /* -1 if no source could supply it                       */
    char     gpu_uuid[48];         /* "GPU-xxxxxxxx-xxxx-xxxx-xxxx-xxxxxxxxxxxx", NUL padded */
    char     pci_bus_id[24];       /* nvidia-smi spelling "00000000:1F:00.0"                */
    char     name[64];
    uint64_t hbm_bytes_total;
    uint32_t sm_count;
    uint32_t cc_major, cc_minor;
    uint32_t identity_source;      /* 1 NVML, 2 /proc, 3 CUDA runtime only                  */
    uint32_t flags;                /* CRO_DEV_*                                             */
    int32_t  dev_index;            /* index for cro_probe_device / cro_hbm_*; -1: not in this process */
    uint32_t reserved[2];
} cro_dev_info;

/* The node's inventory is re-read on EVERY cro_enumerate (the reference execs a fresh nvidia-smi on every
 * reconcile, internal/utils/gpus.go:666-689); the devices a context can probe itself are fixed at cro_probe_init
 * because CUDA's device list is fixed at cuInit. */
#define CRO_DEV_IN_PROCESS    0x1u  /* managed by this context: probe it with cro_probe_device(dev_index)       */
#define CRO_DEV_NEEDS_HELPER  0x2u  /* on the node but attached after cro_probe_init: cro_probe_uuid probes it
                                       through a one-shot helper process (croprobe-cli) with its own cuInit     */

/*
 * Fixed-size, pointer-free, integer-only per-device result: the payload of the
 * single NCCL all-gather.  512 bytes, little-endian.
 * Offsets 0..351 are Appendix C's; its reserved tail holds the rest.
 *
 * The struct is WRITTEN ON THE DEVICE (ABI 2): a one-CTA finalize kernel at the
 * end of the probe's CUDA graph compares every sweep with the closed form,
 * takes the sweep times from %globaltimer and fills the all-gather send buffer;
 * the host only copies it back.  Identity fields are staged once at init.
 *
 * Checksum of a sweep over 64-bit words w[0..n): xor = XOR of all words,
 * sum = wrapping sum, wsum = wrapping sum of w[i] * (2i + 1) — the last one
 * makes every word's POSITION matter.
 */
typedef struct cro_probe_result {
    uint32_t abi_version;          /*   0 */
    int32_t  status;               /*   4  CRO_OK or a CRO_ERR_* */
    int32_t  cuda_ordinal;         /*   8 */
    int32_t  device_minor;         /*  12 */
    char     gpu_uuid[48];         /*  16 */
    char     pci_bus_id[24];       /*  64 */
    uint64_t hbm_bytes_total;      /*  88 */
    uint64_t sweep_bytes;          /*  96 */
    uint64_t seed;                 /* 104  effective pattern seed of THIS probe:
                                           (seed_base | minor) + nonce * 0xD1B54A32D192ED03 */
    uint64_t checksum_xor;         /* 112  of the first read sweep (or of the sweep that failed) */
    uint64_t checksum_sum;         /* 120 */
    uint64_t fill_ns;              /* 128  sweep times: %globaltimer, first CTA start .. last CTA end */
    uint64_t read_best_ns;         /* 136 */
    uint64_t read_median_ns;       /* 144 */
    uint64_t copy_best_ns;         /* 152 */
    uint64_t copy_median_ns;       /* 160 */
    uint32_t sm_count;             /* 168 */
    uint32_t sm_clock_mhz;         /* 172 */
    uint32_t mem_clock_mhz;        /* 176 */
    uint32_t ecc_errors;           /* 180  uncorrected volatile ECC errors (NVML), read at init, at
                                           cro_probe_all and after any failed probe */
    uint64_t p2p_read_ns[8];       /* 184  time to read p2p_bytes out of peer j's HBM over NVLink */
    uint64_t p2p_checksum_xor[8];  /* 248 */
    uint32_t p2p_latency_ns_x16[8];/* 312  mean hop latency x16 (fixed point)        */
    uint8_t  p2p_access[8];        /* 344  cudaDeviceCanAccessPeer                   */
    uint64_t p2p_bytes;            /* 352 */
    uint64_t expect_xor;           /* 360  closed-form checksum computed on the device
                                           by an independent generator kernel        */
    uint64_t expect_sum;           /* 368 */
    uint64_t expect_wsum;          /* 376 */
    uint64_t checksum_wsum;        /* 384 */
    uint64_t copy_checksum_xor;    /* 392  checksum of the LAST copy's destination (read back by the first read sweep) */
    uint64_t copy_checksum_sum;    /* 400 */
    uint64_t copy_checksum_wsum;   /* 408 */
    uint64_t total_ns;             /* 416  first CTA of the fill .. last CTA of the last sweep */
    uint64_t p2p_write_ns[8];      /* 424  time to PUSH p2p_bytes into peer j's scratch half
                                           (posted NVLink writes; the peer re-reads and checks them) */
    uint32_t nonce;                /* 488  probe number on this device (0 = first probe of the context) */
    uint8_t  rank;                 /* 492  index in the minor-sorted device list      */
    uint8_t  world;                /* 493 */
    uint8_t  read_variant;         /* 494  CRO_READ_* actually used                  */
    uint8_t  copy_variant;         /* 495 */
    uint8_t  read_sweeps;          /* 496 */
    uint8_t  copy_sweeps;          /* 497 */
    uint8_t  copy_verified;        /* 498  copy sweeps whose destination was re-read and matched the closed form
                                           (sweep k+1 folds what sweep k wrote; the first read sweep folds the last) */
    uint8_t  fail_code;            /* 499  CRO_FAIL_*: which check failed first (0 = none) */
    uint8_t  fail_index;           /* 500  sweep number / peer index of that check */
    uint8_t  p2p_ok;               /* 501  bit j: NVLink read of, push into and chase through peer j all verified */
    uint8_t  reserved8[2];         /* 502 */
    uint64_t t_start_ns;           /* 504  %globaltimer when the probe's first CTA started */
} cro_probe_result;

#define CRO_FAIL_NONE        0
#define CRO_FAIL_EXPECT      1   /* the closed-form slot is stale or short: the generator kernel did not run */
#define CRO_FAIL_COPY_SRC    2   /* copy sweep fail_index read something else than the pattern */
#define CRO_FAIL_READ        3   /* read sweep fail_index */
#define CRO_FAIL_P2P_READ    4   /* NVLink read of peer fail_index */
#define CRO_FAIL_P2P_PUSH    5   /* what peer fail_index pushed did not land here intact */
#define CRO_FAIL_P2P_CHASE   6   /* pointer chase through peer fail_index ended on the wrong slot */
#define CRO_FAIL_STALE       7   /* a sweep slot carries another probe's nonce: that kernel did not run */

/* CUDA-event times of the sweeps of the device's most recent probe, in launch order
 * (fill, copy sweeps, read sweeps).  kind: 0 fill, 1 copy, 2 read. */
typedef struct cro_sweep_time {
    uint32_t kind;
    uint32_t index;
    uint64_t bytes;                /* algorithmic bytes of the sweep (S, or 2S for a copy) */
    uint64_t event_ns;             /* CUDA events on the launching stream                 */
    uint64_t timer_ns;             /* %globaltimer window the kernel itself recorded      */
} cro_sweep_time;

/* One directed NVLink pair of the most recent cro_probe_all, in full (the 512-byte struct keeps a digest). */
typedef struct cro_p2p_detail {
    uint64_t read_ns, push_ns, reread_ns;
    uint64_t read_xor, read_sum, read_wsum;        /* what this device folded out of the peer's HBM   */
    uint64_t landed_xor, landed_sum, landed_wsum;  /* what the PEER found in its scratch half after this device's push */
    uint64_t expect_xor, expect_sum, expect_wsum;  /* closed form of the peer's first p2p_bytes       */
    uint64_t chase_ns;
    uint32_t chase_end, chase_expect, hops, access;
} cro_p2p_detail;

/* Result of one timed sweep (bench / parity entry points). */
typedef struct cro_sweep_result {
    uint64_t bytes;                /* algorithmic bytes of the sweep (S, or 2S for copy) */
    uint64_t ns;                   /* CUDA-event duration of the kernel launch(es)       */
    uint64_t checksum_xor;
    uint64_t checksum_sum;
    uint32_t variant;
    uint32_t launches;             /* kernels launched by this call                      */
    uint64_t checksum_wsum;        /* position-weighted sum: wrapping sum of w[i] * (2i + 1) */
    uint64_t timer_ns;             /* %globaltimer window the (last) kernel recorded itself */
} cro_sweep_result;

/* ---- lifecycle ----------------------------------------------------------- */

/* Creates the long-lived probe context (CUDA primary contexts, sweep buffers,
 * streams, events).  The reference has no counterpart: it re-execs nvidia-smi
 * every reconcile (internal/utils/gpus.go:886).  A warm context is what makes
 * the storm / churn configs meaningful (SURVEY.md §7 "cold-start cost"). */
int  cro_probe_init(const cro_opts *opts, cro_ctx **out);
void cro_probe_destroy(cro_ctx *ctx);

/* ---- enumeration: replaces the exec of nvidia-smi ------------------------ */

/* Replaces `nvidia-smi --query-gpu=gpu_uuid` run for its side effect by
 * utils.RunNvidiaSmi (internal/utils/gpus.go:666-689). */
int  cro_device_count(cro_ctx *ctx, int *n);

/* Replaces getGPUInfoFromNvidiaPod / getGPUInfoFromCroNodeAgentPod /
 * getGPUInfoFromProcInCroNodeAgentPod (internal/utils/gpus.go:878-919,
 * 921-962, 1014-1089).  Rows are in nvidia-smi order.  The answer is FRESH on every call: the driver's
 * registry under /proc/driver/nvidia/gpus is re-read (and NVML re-initialised when that shows a change), so a GPU
 * composed after cro_probe_init is listed (CRO_DEV_NEEDS_HELPER) and a GPU drained off the bus is not. */
int  cro_enumerate(cro_ctx *ctx, cro_dev_info *out, int cap, int *n);

/* The same merge without a context (ctx-free, no CUDA call): what cro_enumerate answers for a context managing
 * `in_process` on a node whose /proc is mounted at proc_root (NULL = "/proc").  Only /proc is consulted. */
int  cro_node_inventory(const char *proc_root, const cro_dev_info *in_process, int n_in_process,
                        cro_dev_info *out, int cap, int *n);

/* Probe by UUID — the form the reconcile step uses (Status.DeviceID is a UUID,
 * internal/controller/composableresource_controller.go:231-233).  In-process devices: cro_probe_device.  Devices
 * that reached the node after cro_probe_init: a helper process (`croprobe-cli probe-raw`, CUDA_VISIBLE_DEVICES=<uuid>,
 * deadline CRO_HELPER_TIMEOUT_MS).  A UUID the node does not list: CRO_ERR_NO_DEVICE — the reference's
 * "found = false", not an error of the probe.  ctx may be NULL (every device then goes through the helper). */
int  cro_probe_uuid(cro_ctx *ctx, const char *gpu_uuid, cro_probe_result *out);

/* Text that `nvidia-smi --query-gpu=<query> --format=csv,noheader,nounits`
 * would print for these devices, so the unchanged Go parser at
 * internal/utils/gpus.go:903-916 can consume it.  n == 0 emits
 * "No devices were found\n" (gpus.go:896).  query fields: device_minor,
 * gpu_uuid, pci.bus_id, name, index, memory.total. */
int  cro_emit_csv(const cro_dev_info *devs, int n, const char *query,
                  char *buf, size_t cap, size_t *len);

/* The reference's CSV parse rule (internal/utils/gpus.go:896-916), kept
 * quirk-for-quirk.  Output: Go encoding/json of the []map[string]string the
 * reference would build.  exec_err NULL means a nil error.  Returns
 * CRO_ERR_EXEC with the reference's formatted message in buf on the error
 * path, CRO_ERR_PARSE where the reference would panic (short row). */
int  cro_parse_gpu_csv(const char *std_out, const char *std_err, const char *exec_err,
                       const char *query, char *buf, size_t cap, size_t *len);

/* /proc flavour (internal/utils/gpus.go:1045-1089): input is the script's
 * "minor,uuid,bus" lines. */
int  cro_parse_proc_csv(const char *std_out, const char *std_err, const char *exec_err,
                        const char *query, char *buf, size_t cap, size_t *len);

/* Turns one /proc/driver/nvidia/gpus/<bus>/information text into the
 * "minor,uuid,bus" line of the awk script at gpus.go:1017-1037 ("" if any
 * of the three keys is missing). */
int  cro_proc_information_to_line(const char *information_text,
                                  char *buf, size_t cap, size_t *len);

/* Membership decision of utils.CheckGPUVisible, DEVICE_PLUGIN branch
 * (internal/utils/gpus.go:73-84): is device_id among the enumerated UUIDs. */
int  cro_check_gpu_visible(const cro_dev_info *devs, int n, const char *device_id,
                           int *visible);

/* Bus-id / device-path spellings the reference derives
 * (internal/utils/gpus.go:218,326,406,567,238,480).  kind:
 * 0 upper(trim), 1 lower(trim), 2 TrimPrefix("0000") of upper,
 * 3 "/dev/nvidia"+minor, 4 "/run/nvidia/driver/dev/nvidia"+minor. */
int  cro_normalize(int kind, const char *in, char *buf, size_t cap, size_t *len);

/* ---- the probe: new work in the slot of RunNvidiaSmi + CheckGPUVisible --- */

/* Full per-device probe (fill, read sweeps, copy sweeps, closed-form check).
 * Sits at internal/controller/composableresource_controller.go:259 and feeds
 * the decision at :275.  dev_index indexes the cro_enumerate order. */
int  cro_probe_device(cro_ctx *ctx, int dev_index, cro_probe_result *out);

/* Asynchronous form of cro_probe_device: begin enqueues the whole probe on the
 * device's stream and returns at once; end waits for the OLDEST probe begun on
 * the device and hands out its result.  One host thread (the reference's single
 * reconcile worker, MaxConcurrentReconciles=1) can keep every attached GPU busy
 * this way.  Up to TWO probes per device may be in flight: the second one's
 * kernels are queued behind the first's on the device, so the GPU does not idle
 * while the host collects a result and starts the next attach's probe.  A third
 * begin is a no-op; end without begin probes synchronously. */
int  cro_probe_begin(cro_ctx *ctx, int dev_index);
int  cro_probe_end(cro_ctx *ctx, int dev_index, cro_probe_result *out);

/* Concurrent probe of every managed device, NVLink P2P rounds, then ONE
 * ncclAllGather of the 512-byte result structs.  out[] receives the gathered
 * array as rank 0 holds it (asserted byte-identical on every rank).
 * NCCL is dlopen'ed at the first call: the copy the host process already carries if there is one, else
 * $CRO_NCCL_PATH, else libnccl.so.2 — never with RTLD_GLOBAL.  CRO_NCCL_PATH=off, or no usable library: the structs
 * come back per device over pinned memory instead and cro_fullbox_time.gather reports CRO_GATHER_DEGRADED. */
int  cro_probe_all(cro_ctx *ctx, cro_probe_result *out, int cap, int *n);

/* Device address of this device's result struct (the all-gather send buffer),
 * for hosts that run their own collective (bench.py under torchrun). */
int  cro_result_device_ptr(cro_ctx *ctx, int dev_index, uint64_t *dptr);

/* Single sweeps, timed with CUDA events on the launching stream. */
int  cro_hbm_fill(cro_ctx *ctx, int dev_index, cro_sweep_result *out);
int  cro_hbm_read_checksum(cro_ctx *ctx, int dev_index, uint32_t variant, cro_sweep_result *out);
int  cro_hbm_copy(cro_ctx *ctx, int dev_index, uint32_t variant, cro_sweep_result *out);
/* Checksum of the copy destination region (same kernel, other half). */
int  cro_hbm_read_checksum_dst(cro_ctx *ctx, int dev_index, uint32_t variant, cro_sweep_result *out);
/* Closed-form expected checksum computed on the device without touching HBM. */
int  cro_hbm_expected_checksum(cro_ctx *ctx, int dev_index, cro_sweep_result *out);
/* Fault injection for tests: XOR `mask` into the 64-bit word at word_index. */
int  cro_inject_fault(cro_ctx *ctx, int dev_index, uint64_t word_index, uint64_t mask);
/* Copies [word_first, word_first+n_words) of the sweep region to host memory. */
int  cro_read_words(cro_ctx *ctx, int dev_index, uint64_t word_first, uint64_t n_words, uint64_t *out);
/* Repeats the read sweep `iters` times back to back under one event pair. */
int  cro_hbm_read_loop(cro_ctx *ctx, int dev_index, uint32_t variant, uint32_t iters, cro_sweep_result *out);
int  cro_hbm_copy_loop(cro_ctx *ctx, int dev_index, uint32_t variant, uint32_t iters, cro_sweep_result *out);
int  cro_hbm_fill_loop(cro_ctx *ctx, int dev_index, uint32_t iters, cro_sweep_result *out);
/* Seed of the pattern the device's region holds now: (seed_base | minor) + nonce * 0xD1B54A32D192ED03, where
 * nonce counts the probes this context has run on the device (every probe writes a fresh pattern, so a fill or
 * copy that silently did nothing cannot pass on the previous probe's bytes). */
int  cro_device_seed(cro_ctx *ctx, int dev_index, uint64_t *seed);
/* CUDA-event and %globaltimer times of every sweep of the device's most recent probe, in launch order. */
int  cro_probe_sweep_times(cro_ctx *ctx, int dev_index, cro_sweep_time *out, int cap, int *n);
/* One directed NVLink pair (dev_index -> peer_index) of the most recent cro_probe_all. */
int  cro_p2p_detail_get(cro_ctx *ctx, int dev_index, int peer_index, cro_p2p_detail *out);
/* Phases of the most recent cro_probe_all. */
typedef struct cro_fullbox_time {
    uint64_t enqueue_ns;           /* host time spent enqueueing (no waits inside)                      */
    uint64_t wall_ns;              /* host wall clock of the call                                       */
    uint64_t hbm_ns;               /* slowest device's HBM probe (%globaltimer)                         */
    uint64_t p2p_ns;               /* slowest device's NVLink bandwidth rounds, first kernel .. last    */
    uint64_t chase_ns;             /* slowest pointer chase                                             */
    uint64_t gather_ns;            /* the all-gather, CUDA events on rank 0's stream                    */
    uint32_t rounds;               /* NVLink rounds (n-1 for even n)                                    */
    uint32_t host_syncs;           /* stream synchronisations the call made (one per device)            */
    uint32_t gather;               /* CRO_GATHER_*: how the result structs were brought together        */
    uint32_t reserved;
} cro_fullbox_time;
#define CRO_GATHER_HOST      0u    /* copied back per device, assembled on the host (one device, or CRO_F_SKIP_NCCL) */
#define CRO_GATHER_NCCL      1u    /* ncclAllGather on the devices' streams, every rank holds the same array         */
#define CRO_GATHER_DEGRADED  2u    /* NCCL was asked for but no usable libnccl is in reach: host-side gather over
                                      pinned memory instead — "replicas only" (SURVEY.md §8e); cro_last_error says why */
int  cro_fullbox_times(cro_ctx *ctx, cro_fullbox_time *out);
/* The reply structs the fabric decoders walk ("FMScaleUpResponse", "FMGetMachineResponse", "CMMachineData"), as
 * JSON: {"type","struct","fields":[{"json","of":{...}}]} in declaration order — so that a test can hold them against
 * the declarations in the reference's Go source (the .go files of internal/cdi/fti/fm/api, and internal/cdi/fti/cm/api/machine.go). */
int  cro_describe_wire_type(const char *name, char *buf, size_t cap, size_t *len);
/* Prometheus text exposition (counters and per-GPU gauges of the last probe) for the operator's metrics registry
 * (cmd/main.go:66,119-125 wires controller-runtime's registry; a Go collector forwards these lines):
 * cro_probe_total, cro_probe_failures_total, cro_fullbox_probe_total, cro_helper_probe_total,
 * cro_inventory_rescans_total, cro_kernel_launches_total, cro_probe_status{gpu_uuid,minor},
 * cro_probe_hbm_{read,copy,fill}_bytes_per_second{..}, cro_probe_copies_verified{..}, cro_probe_ecc_uncorrected{..}. */
int  cro_metrics_text(cro_ctx *ctx, char *buf, size_t cap, size_t *len);
/* Pointer-chase length of the following cro_probe_all calls (1 .. 16777216 hops per directed pair). */
int  cro_set_latency_hops(cro_ctx *ctx, uint32_t hops);
/* Where `hops` steps from slot 0 of the latency permutation of the directed pair (minor_src chases through
 * minor_dst's memory) end: Sattolo cycle over 65536 slots, mt19937_64 seeded with minor_src * 8 + minor_dst
 * (SURVEY.md §8d config 3).  Host arithmetic only. */
int  cro_chase_end(int minor_src, int minor_dst, uint32_t hops, uint32_t *end);
/* The CRO_* environment knobs are validated like the reference validates its own
 * (internal/controller/composableresource_adapter.go:42-45): an illegal value fails cro_probe_init with
 * "the env variable <NAME> has an invalid value: '<v>'".  This checks one (name, value) pair — or, with
 * name == NULL, the process environment as cro_probe_init would — without needing a GPU. */
int  cro_validate_env(const char *name, const char *value, char *err_buf, size_t err_cap);
/* Kernel launches issued by this context so far (bench "gpu_launches"). */
uint64_t cro_launch_count(cro_ctx *ctx);

/* ---- emit: encoding/json-compatible writers ------------------------------ */

/* ComposableResourceStatus (api/v1alpha1/composableresource_types.go:36-41):
 * {"state":..,"error":..,"device_id":..,"cdi_device_id":..} with Go omitempty
 * rules; bytes identical to json.Marshal. */
int  cro_emit_status_json(const char *state, const char *error, const char *device_id,
                          const char *cdi_device_id, char *buf, size_t cap, size_t *len);

/* ScalarResourceStatus (api/v1alpha1/composabilityrequest_types.go:74-80). */
int  cro_emit_scalar_status_json(const char *state, const char *device_id,
                                 const char *cdi_device_id, const char *node_name,
                                 const char *error, char *buf, size_t cap, size_t *len);

/* FM ScaleUpBody / ScaleDownBody (internal/cdi/fti/fm/api/scale_up.go:19-41,
 * scale_down.go:19-41; built at fti/fm/client.go:115-144, :240-271). */
int  cro_emit_fm_scale_up(const char *tenant_uuid, const char *mach_uuid, const char *res_type,
                          const char *model, char *buf, size_t cap, size_t *len);
int  cro_emit_fm_scale_down(const char *tenant_uuid, const char *mach_uuid, const char *res_type,
                            const char *res_uuid, char *buf, size_t cap, size_t *len);
/* CM resize bodies (internal/cdi/fti/cm/client.go:62-79, :133-139, :211-218). */
int  cro_emit_cm_scale_up(const char *spec_uuid, int device_count,
                          char *buf, size_t cap, size_t *len);
int  cro_emit_cm_scale_down(const char *spec_uuid, int device_count, const char *device_id,
                            char *buf, size_t cap, size_t *len);
/* Sunfish CompositionRequest (internal/cdi/sunfish/client.go:48-61, :78). */
int  cro_emit_sunfish_request(const char *name, long long count, const char *proc_type,
                              const char *model, char *buf, size_t cap, size_t *len);

/* Additive probe annotations (cohdi.io/probe-*), a Go-marshalled
 * map[string]string (keys sorted).  Never mixed into the status bytes. */
int  cro_emit_probe_annotations_json(const cro_probe_result *r, char *buf, size_t cap, size_t *len);

/* (deviceID, CDIDeviceID) from an FM ScaleUpResponse body, with the
 * res_op_status gate of internal/cdi/fti/fm/client.go:184-213.  On the error
 * branches the reference's message is written to err_buf. */
int  cro_fm_parse_scale_up_response(const char *body, const char *resource_name,
                                    const char *res_type, const char *model,
                                    char *device_id, size_t device_id_cap,
                                    char *cdi_device_id, size_t cdi_cap,
                                    char *err_buf, size_t err_cap);

/* CM flavour of the ID production: checkAddingResources
 * (internal/cdi/fti/cm/client.go:432-459,485-509) over the GET-machine JSON.
 * existing_device_ids: '\n'-joined Status.DeviceID of every ComposableResource.
 * Outputs: spec_uuid + *device_count for the resize request, or the unused
 * device's (device_id, cdi_device_id).  Returns CRO_ERR_PARSE with the
 * reference's message in err_buf on the ADD_FAILED / malformed branches (the
 * ids are still filled, as the reference returns them with the error). */
int  cro_cm_check_adding_resources(const char *machine_body, const char *existing_device_ids,
                                   const char *res_type, const char *model,
                                   char *spec_uuid, size_t spec_cap, int *device_count,
                                   char *device_id, size_t device_id_cap,
                                   char *cdi_device_id, size_t cdi_cap,
                                   char *err_buf, size_t err_cap);

/* ---- reconcile step: the caller of the hot path -------------------------- */

/*
 * One Reconcile pass of ComposableResourceReconciler for the state in
 * status.state: "" (handleNoneState :176-198), "Attaching" (handleAttachingState
 * :200-287, the hot path, with the CUDA probe in the RunNvidiaSmi /
 * CheckGPUVisible slots), "Online" (:289-318), "Detaching" (:320-407)
 * (all in internal/controller/composableresource_controller.go).
 *
 * in_json:  {"name":..,"spec":{type,model,target_node,force_detach},"labels":{..},
 *            "status":{state,error,device_id,cdi_device_id},
 *            "deleting":bool,
 *            "device_resource_type":"DEVICE_PLUGIN"|"DRA",
 *            "probe":bool,
 *            "provider":{"device_id","cdi_device_id","error","waiting",          AddResource
 *                        "fm_response_body" | "cm_machine_body"+"existing_device_ids",
 *                        "check_resource_error" | "fm_machine_body" | "cm_check_body",   CheckResource
 *                        "remove":{"waiting","error"}},                                   RemoveResource
 *            what the node would have answered, when there is no live context:
 *            "enumeration":{"stdout","stderr","exec_err"}, "enumeration_after_remove":{..},
 *            "resource_slices":[..], "resource_slices_after_remove":[..],
 *            "driver_pod_missing":bool, "daemonset_errors":{"ns/name":"error"},
 *            "load_check":{"stdout","stderr","exec_err","pod_name","driver_enabled"},
 *            "drain":{"error" | "fd_scan":{"stdout","stderr","exec_err"},"rke2":bool},
 *            "create_taint_error","delete_taint_error",
 *            "status_update_failures":{"after":N,"error":".."}   (the API server refuses Status().Update from attempt
 *                   N+1 on: the handler stops where the reference stops; out_json then carries "failed_status_updates"),
 *            an error text that starts "runtime error: " is a Go panic: no status write, reconcile error
 *            "panic: <text> [recovered]" (what controller-runtime's Reconcile wrapper makes of it),
 *            the DaemonSet restart rule (internal/utils/nodes.go:35-76) instead of canned errors:
 *            "daemonsets":{"ns/name":{"desired","ready","current","unavailable","misscheduled",
 *                                     "restarted_at":"<RFC3339>"}}, "now":"<RFC3339>",
 *            the real FM / CM / Sunfish client (csrc/provider.hpp) instead of "provider":
 *            "env":{"DEVICE_RESOURCE_TYPE","CDI_PROVIDER_TYPE","FTI_CDI_API_TYPE",
 *                   "FTI_CDI_TENANT_ID","FTI_CDI_CLUSTER_ID"}   (adapter selection,
 *                   internal/controller/composableresource_adapter.go:39-72),
 *            "fabric":{"http":[{"method","path"|"path_contains","status","body"}..],
 *                      "transport_error","token_error",
 *                      "objects":{"nodes":{name:{"annotations","provider_id"}},
 *                                 "metal3machines":{"ns/name":{"annotations"}},
 *                                 "baremetalhosts":{"ns/name":{"annotations"}},
 *                                 "composable_resource_device_ids":[..],"status_update_error"}}}
 * out_json: {"status":{...},"requeue_after_s":N,"delete_requested":bool,"error":"..",
 *            "status_updates":[...],"probe":{...},
 *            "daemonset_restarts":["ns/name@<stamp>"..], "fabric_requests":[{method,path,query,body}..]}
 * The status object inside out_json is byte-identical to json.Marshal of the
 * reference's ComposableResourceStatus after the same step.
 */
int  cro_reconcile_attach(cro_ctx *ctx, const char *in_json,
                          char *buf, size_t cap, size_t *len);

/* ---- fabric wire codec (response side) ------------------------------------ */

/* CdiProvider.CheckResource decision over the GET-machine body.  kind "fm":
 * internal/cdi/fti/fm/client.go:314-359; kind "cm": internal/cdi/fti/cm/client.go:262-304.
 * CRO_OK = healthy; CRO_ERR_EXEC = the reference's error text in err_buf. */
int  cro_fabric_check_resource(const char *kind, const char *machine_body, const char *res_type,
                               const char *model, const char *device_id, char *err_buf, size_t err_cap);
/* CdiProvider.GetResources decode for one node's machine (fm/client.go:385-410,
 * cm/client.go:335-343): JSON array of cdi.DeviceInfo
 * {"node_name","machine_uuid","device_type","model","device_id","cdi_device_id"}. */
int  cro_fabric_get_resources(const char *kind, const char *machine_body, const char *node_name,
                              const char *machine_uuid, char *buf, size_t cap, size_t *len);
/* CdiProvider.GetResources of the whole FM / CM client (internal/cdi/fti/fm/client.go:361-413,
 * internal/cdi/fti/cm/client.go:306-346) over a scripted fabric: request_json =
 * {"env": {"CDI_PROVIDER_TYPE","FTI_CDI_API_TYPE","DEVICE_RESOURCE_TYPE","FTI_CDI_TENANT_ID","FTI_CDI_CLUSTER_ID"},
 *  "fabric": {"http": [...], "objects": {...}, "token_error": "" | "token": {...}}} (the same "fabric" object
 * cro_reconcile_attach takes).  Reply: {"devices": [DeviceInfo...], "error": "", "fabric_requests": [...]};
 * the FM flavour skips nodes that fail, the CM flavour aborts with the first error. */
int  cro_fabric_list_devices(const char *request_json, char *buf, size_t cap, size_t *len);
/* What CachedToken.Token makes of the id_manager's answer (internal/cdi/fti/token.go:96-175): reply_json =
 * {"secret_error": "", "transport_error": "", "status": 200, "body": "<reply body>"}; writes
 * {"error": "<text, without the 'unable to rotate token: ' prefix GetToken adds>", "expiry": <exp claim, unix s>}.
 * The same object under "fabric"."token" makes cro_reconcile_attach / cro_fabric_list_devices run the
 * token cache (reuse while expiry - 30 s > "now") in front of every fabric request; they then also report
 * "token_fetches". */
int  cro_token_from_reply(const char *reply_json, char *buf, size_t cap, size_t *len);

/* ---- detach-side pre-flight (the step on the other side of the path) ------ */

/* Parse + decision of utils.CheckNoGPULoads (internal/utils/gpus.go:145-186)
 * over the output of `nvidia-smi --query-compute-apps=gpu_uuid,process_name`.
 * driver_enabled != 0: OCP branch (any load on the node); 0: RKE2 branch (only
 * loads on target_uuid).  CRO_OK = no load; CRO_ERR_EXEC = the reference's
 * error text in err_buf. */
int  cro_check_no_gpu_loads(const char *std_out, const char *std_err, const char *exec_err,
                            const char *pod_name, const char *node_name, const char *target_uuid,
                            int driver_enabled, char *err_buf, size_t err_cap);
/* checkGPUDrainStatus (internal/utils/gpus.go:964-1012) over `nvidia-smi drain -p <bus> -q`. */
int  cro_check_gpu_drain_status(const char *std_out, const char *std_err, const char *exec_err,
                                const char *node_name, const char *bus_id, int *draining,
                                char *err_buf, size_t err_cap);
/* Decision after the open-file scan (gpus.go:468-473, :629-634; RKE2 flavour :286-291). */
int  cro_check_device_file_scan(const char *std_out, const char *std_err, const char *exec_err, int rke2,
                                char *err_buf, size_t err_cap);
/* Native replacement of the fd-scan shell scripts (gpus.go:236-260, 441-457):
 * prints what the script would print for `target` (e.g. "/dev/nvidia0").
 * proc_root NULL = "/proc". */
int  cro_scan_device_file_holders(const char *proc_root, const char *target, int rke2,
                                  char *buf, size_t cap, size_t *len);

/* ---- the caller of the hot path: both reconcilers over an in-memory API --- */

/*
 * In-memory cluster (nodes, ComposabilityRequests, ComposableResources with
 * Kubernetes finalizer / deletionTimestamp semantics) driven by restatements of
 * ComposabilityRequestReconciler (internal/controller/composabilityrequest_controller.go:72-625)
 * and ComposableResourceReconciler (internal/controller/composableresource_controller.go:73-441).
 * Drives BASELINE configs 4 (reconcile storm) and 5 (attach/detach churn);
 * every attach runs the CUDA probe when config.probe is true.
 * config_json: {"nodes":["worker-0",...] | [{"name","cpu","memory","ephemeral_storage","pods"}],
 *               "device_resource_type":"DEVICE_PLUGIN"|"DRA","probe":bool,"seed":N,"uuids":[...]}
 */
typedef struct cro_sim cro_sim;
int  cro_sim_create(cro_ctx *ctx /* may be NULL when probe is false */, const char *config_json, cro_sim **out);
void cro_sim_destroy(cro_sim *sim);
/* kubectl apply: {"name":..,"resource":{type,model,size,force_detach,allocation_policy,target_node,other_spec}} */
int  cro_sim_apply(cro_sim *sim, const char *request_json, char *err_buf, size_t err_cap);
int  cro_sim_delete(cro_sim *sim, const char *request_name);
/* test hook: place an object in a given state ({"kind":"ComposabilityRequest"|"ComposableResource",...}) */
int  cro_sim_plant(cro_sim *sim, const char *object_json, char *err_buf, size_t err_cap);
/* run both controllers until quiescent (or max_reconciles); stats JSON out */
int  cro_sim_run(cro_sim *sim, long long max_reconciles, char *buf, size_t cap, size_t *len);
/* exactly one Reconcile of the request controller (how the reference's tests drive it) */
int  cro_sim_reconcile_request(cro_sim *sim, const char *name, char *err_buf, size_t err_cap);
int  cro_sim_reconcile_resource(cro_sim *sim, const char *name, char *err_buf, size_t err_cap);
/* One tick of the UpstreamSyncer (internal/controller/upstreamsyncer_controller.go:77-159) at time
 * now_s.  devices_json: [{"node_name","machine_uuid","device_type","model","device_id","cdi_device_id"}]
 * (cdi.DeviceInfo, internal/cdi/client.go:25-32) — from the fabric, or from the gathered probe
 * results: a device the node can probe but no ComposableResource owns is the drift it repairs. */
int  cro_sim_sync_upstream(cro_sim *sim, const char *devices_json, long long now_s, char *err_buf, size_t err_cap);
int  cro_sim_dump(cro_sim *sim, char *buf, size_t cap, size_t *len);

/* ---- node-side operations run ON the node ---------------------------------- */

/*
 * One of the node-side operations of internal/utils/gpus.go, executed locally instead of through pod
 * execs: scans are native /proc walks, `nvidia-smi --query-gpu=...` is answered from ctx's enumeration
 * (ctx may be NULL: then nvidia-smi is spawned), every other command is spawned.
 * request_json: {"op": "check_no_gpu_loads" (gpus.go:88-186) | "run_nvidia_smi" (:666-689) |
 *                      "check_gpu_visible" (:54-86) | "drain" (:188-664),
 *                "node": "...", "device_id": "GPU-...", "device_resource_type": "DEVICE_PLUGIN"|"DRA",
 *                "driver_container": bool   (true: the gpu-operator flavour; false: the RKE2 / host-driver flavour),
 *                "allow_mutation": bool     (false = dry run: persistence-mode / drain / rm / modprobe / sysfs
 *                                            writes are logged as skipped and succeed),
 *                "proc_root": "/proc"}
 * Reply: {"error": "<the reference's error text or empty>", "visible": bool,
 *         "exec_log": [{"kind","argv","how": "native"|"spawned"|"skipped (dry run)","failed"}..]}
 */
int  cro_local_node_op(cro_ctx *ctx, const char *request_json, char *buf, size_t cap, size_t *len);
/* One command through the same local executor cro_local_node_op uses, for hosts that drive the flows themselves:
 * request_json = {"argv": [...], "allow_mutation": bool, "exec_deadline_ms": N}.  While allow_mutation is false only
 * the argv shapes known to READ are executed (nvidia-smi --query-gpu / --query-compute-apps, `nvidia-smi drain -p <bus>
 * -q`, lsmod, each optionally behind `/bin/chroot /host-root`); everything else is reported "skipped (dry run)".  A
 * child that outlives exec_deadline_ms (default 60 s) is killed and reaped: "context deadline exceeded".
 * The detach side's nvidia-smi invocations — --query-compute-apps=gpu_uuid,process_name (gpus.go:125,134), drain -p <bus>
 * -q (:970), and with allow_mutation: -i <uuid> -pm 0|1 (:267), drain -p <bus> -m 0|1 (:269), drain -p <bus> -r (:311) —
 * are answered through NVML inside this process when libnvidia-ml is there ("how": "native": no child process, no
 * second NVML init), with nvidia-smi's stdout and exit code (the two queries are pinned byte for byte against the real
 * nvidia-smi on a B200 box; the three mutating texts are not — the reference never parses them).  "native_nvml": false
 * spawns instead; "nvml_lib": "<path>" names another libnvidia-ml (tests).  Both keys work in cro_local_node_op too.
 * Reply: {"how": "spawned"|"skipped (dry run)"|"native", "failed": bool, "exec_err", "stdout", "stderr"}. */
int  cro_local_exec(const char *request_json, char *buf, size_t cap, size_t *len);
/* The cmdline scan of checkResetGPUCommandStillRunning (gpus.go:1182-1226), natively:
 * *found = 1 if another process's command line mentions `needle`. */
int  cro_scan_cmdline_for(const char *proc_root, const char *needle, int *found);

/* ---- diagnostics --------------------------------------------------------- */
const char *cro_strerror(int code);
/* Last error text recorded on this context by the calling thread's most
 * recent failing call (thread-safe copy-out).  ctx == NULL: the text of the
 * calling thread's last failed cro_probe_init (which returns no context), e.g.
 * the cudaMalloc that could not be satisfied. */
int  cro_last_error(cro_ctx *ctx, char *buf, size_t cap);
/* No C++ exception crosses this ABI: every entry point catches, returns CRO_ERR_OOM (std::bad_alloc) or
 * CRO_ERR_INTERNAL (anything else) and keeps the text for cro_last_error(NULL, ...).  This self-test throws on purpose
 * behind the same barrier — kind 0: std::runtime_error, 1: std::bad_alloc, 2: a non-std exception — and returns what the
 * barrier made of it; any other kind returns CRO_OK without throwing. */
int  cro_selftest_exception_barrier(int kind);
const char *cro_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CROPROBE_H_ */
