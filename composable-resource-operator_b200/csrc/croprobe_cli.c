/*
 * croprobe-cli — libcroprobe as a one-shot helper process.
 *
 * Why a helper: CUDA fixes its device list at cuInit, so a long-lived operator process cannot see a GPU that is
 * hot-plugged later (SURVEY.md §7 "freshly hot-plugged GPU"); the reference gets around the same problem by
 * exec'ing nvidia-smi in a pod (internal/utils/gpus.go:886).  A fresh process per attach pays the cold start
 * (cuInit + context + 2*S cudaMalloc) that a warm context avoids; `cold` below prints both so the trade is visible.
 *
 *   croprobe-cli csv <query>            what `nvidia-smi --query-gpu=<query> --format=csv,noheader,nounits` prints
 *   croprobe-cli enumerate              JSON array of the devices (minor, uuid, bus id, name)
 *   croprobe-cli probe <uuid|index> [sweep_MiB]   one full probe; JSON annotations on stdout, exit 0 iff status ok
 *   croprobe-cli cold <index> [sweep_MiB]         timings: init, first (cold) probe, second (warm) probe
 *
 * Plain C against include/croprobe.h — the same surface the cgo shim binds.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "croprobe.h"

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int fail(cro_ctx *ctx, const char *what, int rc) {
    char msg[1024] = {0};
    if (ctx) cro_last_error(ctx, msg, sizeof msg);
    fprintf(stderr, "croprobe-cli: %s: %s%s%s\n", what, cro_strerror(rc), msg[0] ? ": " : "", msg);
    return 2;
}

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: croprobe-cli csv <query> | enumerate | probe <uuid|index> [sweep_MiB] | cold <index> [sweep_MiB]\n");
        return 64;
    }
    const char *cmd = argv[1];
    const int wants_probe = strcmp(cmd, "probe") == 0 || strcmp(cmd, "cold") == 0;
    cro_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.abi_version = CRO_ABI_VERSION;
    opts.flags = CRO_F_LAZY_ALLOC | CRO_F_DEGRADE_ON_OOM;
    if (wants_probe && argc > 3) opts.sweep_bytes = (uint64_t)strtoull(argv[3], NULL, 10) << 20;

    const double t0 = now_s();
    cro_ctx *ctx = NULL;
    int rc = cro_probe_init(&opts, &ctx);
    if (rc == CRO_ERR_NO_DEVICE && strcmp(cmd, "csv") == 0) {   /* an empty box is not an error for enumeration */
        printf("No devices were found\n");
        return 0;
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
            printf("%s{\"index\":%d,\"device_minor\":%d,\"gpu_uuid\":\"%s\",\"pci.bus_id\":\"%s\",\"name\":\"%s\"}", i ? "," : "", i,
                   devs[i].device_minor, devs[i].gpu_uuid, devs[i].pci_bus_id, devs[i].name);
        printf("]\n");
    } else if (wants_probe) {
        if (argc < 3) return 64;
        int idx = -1;
        for (int i = 0; i < n; ++i)
            if (strcmp(devs[i].gpu_uuid, argv[2]) == 0) idx = i;
        if (idx < 0 && argv[2][0] >= '0' && argv[2][0] <= '9') idx = atoi(argv[2]);
        if (idx < 0 || idx >= n) {
            fprintf(stderr, "croprobe-cli: device '%s' is not visible\n", argv[2]);   /* found=0, like gpus.go:896-898 */
            cro_probe_destroy(ctx);
            return 3;
        }
        cro_probe_result r;
        const double t1 = now_s();
        rc = cro_probe_device(ctx, idx, &r);
        const double t_cold = now_s() - t1;
        if (rc != CRO_OK && rc != CRO_ERR_CHECKSUM) return fail(ctx, "cro_probe_device", rc);
        if (strcmp(cmd, "cold") == 0) {
            const double t2 = now_s();
            cro_probe_result r2;
            int rc2 = cro_probe_device(ctx, idx, &r2);
            const double t_warm = now_s() - t2;
            printf("{\"gpu_uuid\":\"%s\",\"sweep_bytes\":%llu,\"init_s\":%.4f,\"cold_probe_s\":%.4f,\"warm_probe_s\":%.4f,"
                   "\"cold_total_s\":%.4f,\"cold_probes_per_s\":%.2f,\"warm_probes_per_s\":%.2f,\"status\":%d}\n",
                   r.gpu_uuid, (unsigned long long)r.sweep_bytes, t_init, t_cold, t_warm, t_init + t_cold,
                   1.0 / (t_init + t_cold), 1.0 / t_warm, rc2 ? rc2 : r.status);
        } else {
            if ((rc = cro_emit_probe_annotations_json(&r, buf, sizeof buf, &len)) != CRO_OK) return fail(ctx, "emit", rc);
            printf("%s\n", buf);
        }
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
