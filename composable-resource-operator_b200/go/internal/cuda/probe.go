// Package cuda is the thin cgo shim over libcroprobe (include/croprobe.h).
//
// UNCOMPILED IN THIS REPOSITORY'S BUILD IMAGE: there is no Go toolchain there
// (`go version`: command not found).  Every behaviour claimed for this file is
// exercised through the same C entry points by the ctypes tests in tests/.
//
// It is a new sibling of internal/utils (SURVEY.md §1, layer L1c) and is called
// from handleAttachingState at the two slots the reference fills with
// utils.RunNvidiaSmi (internal/controller/composableresource_controller.go:259)
// and utils.CheckGPUVisible (:275).
package cuda

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -lcroprobe -ldl -lpthread
#include <stdlib.h>
#include "croprobe.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// ProbeResult mirrors cro_probe_result (512 bytes, integer-only).
type ProbeResult struct {
	Status       int32
	CudaOrdinal  int32
	DeviceMinor  int32
	GPUUUID      string
	PCIBusID     string
	SweepBytes   uint64
	ChecksumXor  uint64
	ChecksumSum  uint64
	ChecksumWsum uint64 // position-weighted sum: every word's place matters
	Nonce        uint32 // probe number on this device; every probe writes a fresh pattern
	CopyVerified uint8  // copy sweeps whose destination was re-read and matched the closed form
	FailCode     uint8  // CRO_FAIL_*: which device-side check failed first
	FailIndex    uint8
	P2POk        uint8 // bit j: NVLink read / push / chase through peer j verified
	FillNs       uint64
	ReadBestNs   uint64
	CopyBestNs   uint64
	P2PReadNs    [8]uint64
	P2PWriteNs   [8]uint64
	P2PLatencyNs [8]uint32
	EccErrors    uint32 // uncorrected volatile ECC errors (NVML), read at init / full-box probe / failed probe
	Annotations  string // Go-marshalled map[string]string of cohdi.io/probe-* keys
}

// OK mirrors `probe.OK` in SURVEY.md §3.3.
func (r ProbeResult) OK() bool { return r.Status == C.CRO_OK }

// Options are the env-style knobs (SURVEY.md §5: env vars with validated values).
type Options struct {
	SweepBytes uint64 // 0 = 4 GiB
	DeadlineMs int32  // a Go ctx cannot cross cgo; pass its remaining time here
	Flags      uint32
}

// Context is the long-lived probe context.  Create it once per manager
// process; it is safe for concurrent use (per-device mutexes inside the
// library, cudaSetDevice in every entry point, so goroutine migration between
// OS threads is harmless).
type Context struct {
	mu sync.Mutex
	h  *C.cro_ctx
}

func errorOf(h *C.cro_ctx, rc C.int) error {
	if rc == C.CRO_OK {
		return nil
	}
	buf := (*C.char)(C.malloc(1024))
	defer C.free(unsafe.Pointer(buf))
	detail := ""
	if h != nil && C.cro_last_error(h, buf, 1024) == C.CRO_OK {
		detail = ": " + C.GoString(buf)
	}
	// stable prefix, surfaced verbatim into Status.Error by requeueOnErr
	// (internal/controller/composableresource_controller.go:423-433)
	return fmt.Errorf("cuda probe failed: %s%s", C.GoString(C.cro_strerror(rc)), detail)
}

// NewContext wraps cro_probe_init.
func NewContext(o Options) (*Context, error) {
	var opts C.cro_opts
	opts.abi_version = C.CRO_ABI_VERSION
	opts.sweep_bytes = C.uint64_t(o.SweepBytes)
	opts.deadline_ms = C.int32_t(o.DeadlineMs)
	opts.flags = C.uint32_t(o.Flags)
	c := &Context{}
	if rc := C.cro_probe_init(&opts, &c.h); rc != C.CRO_OK {
		return nil, errorOf(nil, rc)
	}
	runtime.SetFinalizer(c, func(c *Context) { c.Close() })
	return c, nil
}

// Close wraps cro_probe_destroy.
func (c *Context) Close() {
	c.mu.Lock()
	defer c.mu.Unlock()
	if c.h != nil {
		C.cro_probe_destroy(c.h)
		c.h = nil
	}
}

// EnumerateCSV returns the text `nvidia-smi --query-gpu=<query>
// --format=csv,noheader,nounits` would print, so the UNCHANGED parser at
// internal/utils/gpus.go:903-916 can consume it ("No devices were found" for
// an empty box, gpus.go:896).
func (c *Context) EnumerateCSV(query string) (string, error) {
	var devs [C.CRO_MAX_DEVICES]C.cro_dev_info
	var n C.int
	if rc := C.cro_enumerate(c.h, &devs[0], C.CRO_MAX_DEVICES, &n); rc != C.CRO_OK {
		return "", errorOf(c.h, rc)
	}
	q := C.CString(query)
	defer C.free(unsafe.Pointer(q))
	buf := (*C.char)(C.malloc(8192)) // caller-allocated; C does not retain it
	defer C.free(unsafe.Pointer(buf))
	var ln C.size_t
	if rc := C.cro_emit_csv(&devs[0], n, q, buf, 8192, &ln); rc != C.CRO_OK {
		return "", errorOf(c.h, rc)
	}
	return C.GoStringN(buf, C.int(ln)), nil
}

func convert(r *C.cro_probe_result) ProbeResult {
	out := ProbeResult{
		Status: int32(r.status), CudaOrdinal: int32(r.cuda_ordinal), DeviceMinor: int32(r.device_minor),
		GPUUUID: C.GoString(&r.gpu_uuid[0]), PCIBusID: C.GoString(&r.pci_bus_id[0]),
		SweepBytes: uint64(r.sweep_bytes), ChecksumXor: uint64(r.checksum_xor), ChecksumSum: uint64(r.checksum_sum),
		ChecksumWsum: uint64(r.checksum_wsum), Nonce: uint32(r.nonce), CopyVerified: uint8(r.copy_verified),
		FailCode: uint8(r.fail_code), FailIndex: uint8(r.fail_index), P2POk: uint8(r.p2p_ok),
		FillNs: uint64(r.fill_ns), ReadBestNs: uint64(r.read_best_ns), CopyBestNs: uint64(r.copy_best_ns),
		EccErrors: uint32(r.ecc_errors),
	}
	for j := 0; j < 8; j++ {
		out.P2PReadNs[j] = uint64(r.p2p_read_ns[j])
		out.P2PWriteNs[j] = uint64(r.p2p_write_ns[j])
		out.P2PLatencyNs[j] = uint32(r.p2p_latency_ns_x16[j]) / 16
	}
	buf := (*C.char)(C.malloc(4096))
	defer C.free(unsafe.Pointer(buf))
	var ln C.size_t
	if C.cro_emit_probe_annotations_json(r, buf, 4096, &ln) == C.CRO_OK {
		out.Annotations = C.GoStringN(buf, C.int(ln))
	}
	return out
}

// ProbeAll wraps cro_probe_all: concurrent probe of every attached GPU, NVLink
// rounds, one NCCL all-gather of the result structs.
func (c *Context) ProbeAll() ([]ProbeResult, error) {
	var res [C.CRO_MAX_DEVICES]C.cro_probe_result
	var n C.int
	rc := C.cro_probe_all(c.h, &res[0], C.CRO_MAX_DEVICES, &n)
	if rc != C.CRO_OK && rc != C.CRO_ERR_CHECKSUM {
		return nil, errorOf(c.h, rc)
	}
	out := make([]ProbeResult, 0, int(n))
	for i := 0; i < int(n); i++ {
		out = append(out, convert(&res[i]))
	}
	return out, nil
}

// ProbeUUID probes the one device whose UUID is deviceID (Status.DeviceID,
// internal/controller/composableresource_controller.go:231-233).  The library
// re-reads the node's inventory on every call (the reference execs a fresh
// nvidia-smi per reconcile, internal/utils/gpus.go:666-689): a device this
// context holds is probed in process, a device that reached the node AFTER
// cro_probe_init — which no running CUDA process can see — is probed by a
// one-shot helper process with its own cuInit.  found=false (CRO_ERR_NO_DEVICE)
// mirrors the reference's "not yet visible" (false, nil) + RequeueAfter 30 s.
func (c *Context) ProbeUUID(deviceID string) (r ProbeResult, found bool, err error) {
	id := C.CString(deviceID)
	defer C.free(unsafe.Pointer(id))
	var res C.cro_probe_result
	rc := C.cro_probe_uuid(c.h, id, &res)
	if rc == C.CRO_ERR_NO_DEVICE {
		return r, false, nil
	}
	if rc != C.CRO_OK {
		return convert(&res), true, errorOf(c.h, rc)
	}
	return convert(&res), true, nil
}

// MetricsText is the Prometheus text exposition of the context's counters and
// per-GPU gauges; a prometheus.Collector registered with
// sigs.k8s.io/controller-runtime/pkg/metrics.Registry (cmd/main.go:66,119-125
// wires that registry) forwards it.
func (c *Context) MetricsText() (string, error) {
	buf := (*C.char)(C.malloc(16384))
	defer C.free(unsafe.Pointer(buf))
	var ln C.size_t
	if rc := C.cro_metrics_text(c.h, buf, 16384, &ln); rc != C.CRO_OK {
		return "", errorOf(c.h, rc)
	}
	return C.GoStringN(buf, C.int(ln)), nil
}

// LocalNodeOp runs one node-side operation of internal/utils/gpus.go on the node
// itself (a cro-node-agent that links libcroprobe): request is
// {"op": "check_no_gpu_loads"|"run_nvidia_smi"|"check_gpu_visible"|"drain",
// "node", "device_id", "device_resource_type", "driver_container",
// "allow_mutation"}.  The flows, their exec sequence and their error strings are
// the reference's (CheckNoGPULoads :88-186, DrainGPU :188-664); the /proc scans
// are native and the nvidia-smi steps (compute apps, drain -q / -m 1 / -r,
// -pm 0) are NVML calls in this process, so a detach pre-flight is a handful of
// library calls instead of 3-8 SPDY execs.  The reply's "error" is what the
// reference would have returned ("" = nil); "exec_log" says how each step ran.
func (c *Context) LocalNodeOp(requestJSON string) (string, error) {
	req := C.CString(requestJSON)
	defer C.free(unsafe.Pointer(req))
	buf := (*C.char)(C.malloc(65536))
	defer C.free(unsafe.Pointer(buf))
	var ln C.size_t
	if rc := C.cro_local_node_op(c.h, req, buf, 65536, &ln); rc != C.CRO_OK {
		return "", errorOf(c.h, rc)
	}
	return C.GoStringN(buf, C.int(ln)), nil
}

// Visible is the decision of utils.CheckGPUVisible (internal/utils/gpus.go:73-84)
// strengthened: listed AND the probe reproduced the HBM pattern.
func Visible(results []ProbeResult, deviceID string) bool {
	for _, r := range results {
		if r.GPUUUID == deviceID {
			return r.OK()
		}
	}
	return false
}
