/*
 * croprobe-cli — libcroprobe as a one-shot helper process.
 *
 * Why a helper: CUDA fixes its device list at cuInit, so a long-lived operator process cannot see a GPU that is
 * hot-plugged later (SURVEY.md §7 "freshly hot-plugged GPU"); the reference gets around the same problem by
 * exec'ing nvidia-smi in a pod on every reconcile (internal/utils/gpus.go:886).  libcroprobe keeps the node's
 * inventory fresh by re-reading the driver's registry, and when a device shows up that its CUDA contexts do not
 * cover, cro_probe_uuid runs THIS program for it (`probe-raw`, CUDA_VISIBLE_DEVICES=<uuid>): a fresh cuInit that
 * sees exactly that GPU.  The cold start it pays (cuInit + one primary context + cudaMalloc) is what `cold` prints.
 *
 *   croprobe-cli csv <query>            what `nvidia-smi --query-gpu=<query> --format=csv,noheader,nounits` prints
 *   croprobe-cli enumerate              JSON array of the devices (minor, uuid, bus id, name)
 *   croprobe-cli probe <uuid|index> [sweep_MiB]      one full probe; JSON annotations on stdout, exit 0 iff status ok
 *   croprobe-cli probe-raw <uuid> [sweep_MiB]        the same, the 512-byte cro_probe_result on stdout (for the library)
 *   croprobe-cli cold <uuid|index> [sweep_MiB] [nvml] timings: init, first (cold) probe, second (warm) probe
 * Exit 3: the device is not visible to this (fresh) process — the reference's found=false.
 *
 * Plain C against include/croprobe.h — the same surface the cgo shim binds.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "croprobe.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int fail(cro_ctx *ctx, const char *what, int rc) {
    char msg[1024] = {0};
    cro_last_error(ctx, msg, sizeof msg);
    fprintf(stderr, "croprobe-cli: %s: %s%s%s\n", what, cro_strerror(rc), msg[0] ? ": " : "", msg);
    return 2;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: croprobe-cli csv <query> | enumerate | probe <uuid|index> [sweep_MiB] | probe-raw <uuid> [sweep_MiB] | "
                        "cold <uuid|index> [sweep_MiB] [nvml]\n");
        return 64;
    }
    const double t_start = now_s();
    const char *cmd = argv[1];
    const int raw = strcmp(cmd, "probe-raw") == 0;
    const int cold = strcmp(cmd, "cold") == 0;
    const int wants_probe = raw || cold || strcmp(cmd, "probe") == 0;
    if (wants_probe && argc < 3) return 64;
    cro_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.abi_version = CRO_ABI_VERSION;
    opts.flags = CRO_F_LAZY_ALLOC | CRO_F_DEGRADE_ON_OOM;
    if (wants_probe) {
        /* The hot-plug path: one device, identity from /proc (NVML's first call costs more than the probe), and a
         * 1 GiB first sweep unless told otherwise — it already runs at ~7 TB/s and shortens everything before it. */
        opts.sweep_bytes = (argc > 3 ? (uint64_t)strtoull(argv[3], NULL, 10) : 1024ull) << 20;
        if (!(cold && argc > 4 && strcmp(argv[4], "nvml") == 0)) opts.flags |= CRO_F_NO_NVML;
        if (strncmp(argv[2], "GPU-", 4) == 0) {
            /* before ANY CUDA call: this process's cuInit must enumerate that one GPU only (one primary context
             * instead of eight on a full box) */
            if (!getenv("CUDA_VISIBLE_DEVICES") || strcmp(getenv("CUDA_VISIBLE_DEVICES"), argv[2]) != 0)
                setenv("CUDA_VISIBLE_DEVICES", argv[2], 1);
        } else {
            opts.n_devices = 1;
            opts.devices[0] = atoi(argv[2]);
        }
    }

    const double t0 = now_s();
    cro_ctx *ctx = NULL;
    int rc = cro_probe_init(&opts, &ctx);
    if (rc == CRO_ERR_NO_DEVICE && strcmp(cmd, "csv") == 0) {   /* an empty box is not an error for enumeration */
        printf("No devices were found\n");
        return 0;
    }
    if (rc == CRO_ERR_NO_DEVICE && wants_probe && strncmp(argv[2], "GPU-", 4) == 0) {
        fprintf(stderr, "croprobe-cli: device '%s' is not visible\n", argv[2]);   /* found=0, like gpus.go:896-898 */
        return 3;
    }
    if (rc != CRO_OK) return fail(NULL, "cro_probe_init", rc);
    const double t_init = now_s() - t0;

    cro_dev_info devs[CRO_MAX_DEVICES];
    int n = 0;
    if ((rc = cro_enumerate(ctx, devs, CRO_MAX_DEVICES, &n)) != CRO_OK) return fail(ctx, "cro_enumerate", rc);
    static char buf[1 << 16];
    size_t len = 0;

    if (strcmp(cmd, "csv") == 0) {
        if ((rc = cro_emit_csv(devs, n, argc > 2 ? argv[2] : "gpu_uuid", buf, sizeof buf, &len)) != CRO_OK) {
            fprintf(stdout, "%s\n", buf);   /* nvidia-smi prints its field error on stdout too */
            return 2;
        }
        fputs(buf, stdout);
    } else if (strcmp(cmd, "enumerate") == 0) {
        printf("[");
        for (int i = 0; i < n; ++i)
            printf("%s{\"index\":%d,\"device_minor\":%d,\"gpu_uuid\":\"%s\",\"pci.bus_id\":\"%s\",\"name\":\"%s\",\"in_process\":%s}", i ? "," : "",
                   devs[i].dev_index, devs[i].device_minor, devs[i].gpu_uuid, devs[i].pci_bus_id, devs[i].name,
                   (devs[i].flags & CRO_DEV_IN_PROCESS) ? "true" : "false");
        printf("]\n");
    } else if (wants_probe) {
        int idx = -1;
        for (int i = 0; i < n; ++i)
            if ((devs[i].flags & CRO_DEV_IN_PROCESS) && (strcmp(devs[i].gpu_uuid, argv[2]) == 0 || strncmp(argv[2], "GPU-", 4) != 0))
                idx = devs[i].dev_index;
        if (idx < 0) {
            fprintf(stderr, "croprobe-cli: device '%s' is not visible\n", argv[2]);
            cro_probe_destroy(ctx);
            return 3;
        }
        cro_probe_result r;
        const double t1 = now_s();
        rc = cro_probe_device(ctx, idx, &r);
        const double t_cold = now_s() - t1;
        if (rc != CRO_OK && rc != CRO_ERR_CHECKSUM) return fail(ctx, "cro_probe_device", rc);
        if (cold) {
            const double t2 = now_s();
            cro_probe_result r2;
            int rc2 = cro_probe_device(ctx, idx, &r2);
            const double t_warm = now_s() - t2;
            const double total = now_s() - t_start - t_warm;     /* process start .. first verdict */
            printf("{\"gpu_uuid\":\"%s\",\"sweep_bytes\":%llu,\"init_s\":%.4f,\"cold_probe_s\":%.4f,\"warm_probe_s\":%.4f,"
                   "\"cold_total_s\":%.4f,\"cold_probes_per_s\":%.2f,\"warm_probes_per_s\":%.2f,\"nvml\":%s,\"status\":%d}\n",
                   r.gpu_uuid, (unsigned long long)r.sweep_bytes, t_init, t_cold, t_warm, total,
                   1.0 / total, 1.0 / t_warm, (opts.flags & CRO_F_NO_NVML) ? "false" : "true", rc2 ? rc2 : r.status);
        } else if (raw) {
            if (fwrite(&r, sizeof r, 1, stdout) != 1) return 2;
            fflush(stdout);
        } else {
            if ((rc = cro_emit_probe_annotations_json(&r, buf, sizeof buf, &len)) != CRO_OK) return fail(ctx, "emit", rc);
            printf("%s\n", buf);
        }
        if (raw) _exit(r.status == CRO_OK ? 0 : 1);    /* the verdict is out: skip the teardown (it costs as much as the probe) */
        cro_probe_destroy(ctx);
        return r.status == CRO_OK ? 0 : 1;
    } else {
        fprintf(stderr, "croprobe-cli: unknown command '%s'\n", cmd);
        cro_probe_destroy(ctx);
        return 64;
    }
    cro_probe_destroy(ctx);
    return 0;
}
