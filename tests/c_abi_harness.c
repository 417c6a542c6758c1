/*
 * c_abi_harness.c — calls libcroprobe the way the cgo shim does: plain C, caller-allocated buffers, no C++
 * or torch types anywhere.  Compiled with gcc (not nvcc, not g++) against include/croprobe.h, so it also
 * proves the header is valid C.  Without a GPU it exercises the text entry points and checks that
 * cro_probe_init refuses loudly; with a GPU (argv[1] == "gpu") it runs one probe and one attach reconcile.
 *
 * Build: gcc -std=c11 -Wall -Wextra -Iinclude tests/c_abi_harness.c -o /tmp/c_abi_harness \
 *            -Lcomposable-resource-operator_b200 -lcroprobe -Wl,-rpath,$PWD/composable-resource-operator_b200
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "croprobe.h"

#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);  \
            return 1;                                                        \
        }                                                                    \
    } while (0)

int main(int argc, char **argv) {
    char buf[8192], err[1024];
    size_t len = 0;

    CHECK(sizeof(cro_probe_result) == 512);
    CHECK(strcmp(cro_strerror(CRO_OK), "ok") == 0);

    /* emit: the status JSON of the reference's Online KAT */
    CHECK(cro_emit_status_json("Online", "", "GPU-device00-uuid-temp-0000-000000000000",
                               "GPU-device00-uuid-temp-0000-000000000res", buf, sizeof buf, &len) == CRO_OK);
    CHECK(strcmp(buf, "{\"state\":\"Online\",\"device_id\":\"GPU-device00-uuid-temp-0000-000000000000\","
                      "\"cdi_device_id\":\"GPU-device00-uuid-temp-0000-000000000res\"}") == 0);
    CHECK(len == strlen(buf));

    /* too-small buffer reports the needed size */
    CHECK(cro_emit_status_json("Online", "", "", "", buf, 4, &len) == CRO_ERR_BUFFER_SMALL && len == 18);

    /* parse: the reference's 3-field mock line */
    CHECK(cro_parse_gpu_csv("0, GPU-device00-uuid-temp-0000-000000000000, 00000000:1F:00.0", "", NULL,
                            "device_minor,gpu_uuid,pci.bus_id", buf, sizeof buf, &len) == CRO_OK);
    CHECK(strstr(buf, "\"pci.bus_id\":\"00000000:1F:00.0\"") != NULL);
    CHECK(cro_parse_gpu_csv("", "nvidia-smi: command not found", NULL, "gpu_uuid", buf, sizeof buf, &len) == CRO_ERR_EXEC);
    CHECK(strcmp(buf, "get gpu info command failed: err: '<nil>', stderr: 'nvidia-smi: command not found', stdout: ''") == 0);

    /* attach step without a context: enumeration text injected, probe off */
    const char *req =
        "{\"name\":\"test-composable-resource\",\"spec\":{\"type\":\"gpu\",\"model\":\"NVIDIA-A100-PCIE-80GB\","
        "\"target_node\":\"worker-0\"},\"status\":{\"state\":\"Attaching\"},\"device_resource_type\":\"DEVICE_PLUGIN\","
        "\"probe\":false,\"provider\":{\"device_id\":\"GPU-device00-uuid-temp-0000-000000000000\","
        "\"cdi_device_id\":\"GPU-device00-uuid-temp-0000-000000000res\"},"
        "\"enumeration\":{\"stdout\":\"GPU-device00-uuid-temp-0000-000000000000\",\"stderr\":\"\"}}";
    CHECK(cro_reconcile_attach(NULL, req, buf, sizeof buf, &len) == CRO_OK);
    CHECK(strstr(buf, "\"status\":{\"state\":\"Online\",\"device_id\":\"GPU-device00-uuid-temp-0000-000000000000\"") != NULL);

    /* detach pre-flight */
    CHECK(cro_check_device_file_scan("nvidia-persist", "", NULL, 0, err, sizeof err) == CRO_ERR_EXEC);
    CHECK(strcmp(err, "check /dev/nvidiaX command failed: there is a process nvidia-persist occupied the nvidiaX file") == 0);

    cro_opts opts;
    memset(&opts, 0, sizeof opts);
    opts.abi_version = CRO_ABI_VERSION;
    opts.sweep_bytes = 64ull << 20;
    opts.read_sweeps = 2;
    opts.copy_sweeps = 1;
    opts.n_devices = 1;
    opts.devices[0] = 0;
    cro_ctx *ctx = NULL;
    int rc = cro_probe_init(&opts, &ctx);

    if (argc > 1 && strcmp(argv[1], "gpu") == 0) {
        CHECK(rc == CRO_OK && ctx != NULL);
        cro_dev_info devs[CRO_MAX_DEVICES];
        int n = 0;
        CHECK(cro_enumerate(ctx, devs, CRO_MAX_DEVICES, &n) == CRO_OK && n >= 1);   /* the whole NODE, fresh */
        int mine = -1;
        for (int i = 0; i < n; ++i)
            if (devs[i].flags & CRO_DEV_IN_PROCESS) { CHECK(mine < 0 && devs[i].dev_index == 0); mine = i; }
        CHECK(mine >= 0);
        devs[0] = devs[mine];
        cro_probe_result r;
        CHECK(cro_probe_device(ctx, 0, &r) == CRO_OK);
        CHECK(r.status == CRO_OK && r.checksum_xor == r.expect_xor && r.checksum_sum == r.expect_sum && r.checksum_wsum == r.expect_wsum);
        CHECK(r.abi_version == CRO_ABI_VERSION && r.fail_code == CRO_FAIL_NONE && r.copy_verified == r.copy_sweeps);
        CHECK(strcmp(r.gpu_uuid, devs[0].gpu_uuid) == 0);
        cro_probe_result r2;
        CHECK(cro_probe_uuid(ctx, devs[0].gpu_uuid, &r2) == CRO_OK && r2.nonce == r.nonce + 1);
        CHECK(cro_probe_uuid(ctx, "GPU-00000000-dead-beef-0000-000000000000", &r2) == CRO_ERR_NO_DEVICE);
        CHECK(cro_emit_csv(devs, n, "gpu_uuid", buf, sizeof buf, &len) == CRO_OK);
        int visible = 0;
        CHECK(cro_check_gpu_visible(devs, n, devs[0].gpu_uuid, &visible) == CRO_OK && visible == 1);
        CHECK(cro_emit_probe_annotations_json(&r, buf, sizeof buf, &len) == CRO_OK);
        printf("gpu ok: %s read %.1f GB/s launches %llu\n%s\n", r.gpu_uuid, (double)r.sweep_bytes / (double)r.read_best_ns,
               (unsigned long long)cro_launch_count(ctx), buf);
        cro_probe_destroy(ctx);
    } else if (rc == CRO_OK) {
        /* a GPU happened to be present: fine, just clean up */
        cro_probe_destroy(ctx);
    } else {
        CHECK(ctx == NULL);
        CHECK(rc == CRO_ERR_NO_DEVICE || rc == CRO_ERR_CUDA);   /* no CPU fallback: loud refusal */
    }
    printf("c abi harness ok\n");
    return 0;
}
