// Package b200sched is the cgo shim over libb200sched.so (include/b200sched.h): the device implementation of
// Scheduler.Solve (vendor/sigs.k8s.io/karpenter/pkg/controllers/provisioning/scheduling/scheduler.go:377-432) and of
// what it calls, for gpu-provisioner built with CGO_ENABLED=1 (go/Dockerfile).
//
// NOT COMPILED IN THE AUTHORING IMAGE (no Go toolchain there, SURVEY.md fact 4).  Every call below is mirrored by the
// ctypes binding gpu-provisioner_b200/b200sched/capi.py, which IS exercised by the test suite against the same library;
// flatten.go's rules are mirrored by gpu-provisioner_b200/b200sched/flatten.py (tests/test_flatten.py).
//
// Conventions of the C-ABI (include/b200sched.h): return 0 / negative code, the caller allocates every output buffer,
// C keeps no Go pointer after a call returns (cgo pointer rule; slices are pinned for the duration of the call), one
// CUDA stream per context, a context is not re-entrant (guarded by a mutex here; the provisioner and the disruption
// controller each own one).  Any error makes the caller fall back to the vendored CPU Solve: ErrFallback.
package b200sched

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -lb200sched -lcudart
#include <stdlib.h>
#include "b200sched.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// ErrFallback wraps every failure of the device path: the caller keeps the vendored CPU Solve for this batch.
var ErrFallback = errors.New("b200sched: falling back to the CPU scheduler")

// Context owns one b200s_ctx_t (one device, one stream) and the catalogs uploaded to it.
type Context struct {
	mu  sync.Mutex
	ctx *C.b200s_ctx_t
	// catalogs maps the key layout of an encode (b200s_catalog_soa_t.layout_hash) to the id of the catalog uploaded
	// from it.  The layout depends on the keys the batch's pods use, so a later batch may need another upload
	// (B200S_ESTALE otherwise); the GetInstanceTypes generation is part of the key because a new SKU list is a new catalog.
	catalogs map[catalogKey]uint64
}

type catalogKey struct {
	generation uint64
	layout     uint64
}

// NewContext initialises the library on one sm_100 device.  There is no CPU backend: without a device the error is
// B200S_ENODEVICE and the controller runs the vendored scheduler only.
func NewContext(device int) (*Context, error) {
	var ctx *C.b200s_ctx_t
	if rc := C.b200s_init(C.int(device), &ctx); rc != 0 {
		return nil, fmt.Errorf("%w: b200s_init=%d", ErrFallback, int(rc))
	}
	c := &Context{ctx: ctx, catalogs: map[catalogKey]uint64{}}
	runtime.SetFinalizer(c, func(c *Context) { c.Close() })
	return c, nil
}

// Close releases the device context and every uploaded catalog.
func (c *Context) Close() {
	c.mu.Lock()
	defer c.mu.Unlock()
	if c.ctx != nil {
		C.b200s_destroy(c.ctx)
		c.ctx = nil
	}
}

func (c *Context) lastError() string { return C.GoString(C.b200s_last_error(c.ctx)) }

// encoded is a b200s_encoded_t with its owner.
type encoded struct{ p *C.b200s_encoded_t }

func (e *encoded) free() {
	if e.p != nil {
		C.b200s_encoded_free(e.p)
		e.p = nil
	}
}

// encode runs b200s_encode on a flattened model.  B200S_ENOTREPRESENTABLE (topology spread, minValues, more than 64
// values of a label key, ...) is the ordinary way a batch ends up on the CPU path.
func encode(m *Model, pin *runtime.Pinner) (*encoded, error) {
	cm := m.toC(pin)
	var enc *C.b200s_encoded_t
	var why [256]C.char
	if rc := C.b200s_encode(&cm, &enc, &why[0], 256); rc != 0 {
		return nil, fmt.Errorf("%w: b200s_encode=%d %s", ErrFallback, int(rc), C.GoString(&why[0]))
	}
	return &encoded{enc}, nil
}

// catalogFor returns the id of the catalog that matches enc's key layout, uploading it when this layout (of this
// GetInstanceTypes generation) has not been seen.  Must be called with c.mu held.
func (c *Context) catalogFor(enc *encoded, generation uint64) (uint64, error) {
	cat := C.b200s_encoded_catalog(enc.p)
	key := catalogKey{generation, uint64(cat.layout_hash)}
	if id, ok := c.catalogs[key]; ok {
		return id, nil
	}
	var id C.uint64_t
	if rc := C.b200s_catalog_upload(c.ctx, cat, &id); rc != 0 {
		return 0, fmt.Errorf("%w: b200s_catalog_upload=%d %s", ErrFallback, int(rc), c.lastError())
	}
	for k, old := range c.catalogs { // catalogs of older GetInstanceTypes generations are not needed any more
		if k.generation != generation {
			C.b200s_catalog_release(c.ctx, C.uint64_t(old))
			delete(c.catalogs, k)
		}
	}
	c.catalogs[key] = uint64(id)
	return uint64(id), nil
}

// Solve runs one Scheduler.Solve on the device: flatten (the caller) -> encode -> upload if needed -> solve -> decode.
// generation identifies the GetInstanceTypes result the model was flattened from.
func (c *Context) Solve(m *Model, generation uint64) (*Result, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	if c.ctx == nil {
		return nil, fmt.Errorf("%w: context closed", ErrFallback)
	}
	var pin runtime.Pinner
	defer pin.Unpin()
	enc, err := encode(m, &pin)
	if err != nil {
		return nil, err
	}
	defer enc.free()
	id, err := c.catalogFor(enc, generation)
	if err != nil {
		return nil, err
	}
	res := newResult(m, enc)
	res.pin(&pin)
	var nodes *C.b200s_nodes_soa_t
	if m.NumNodes() > 0 {
		nodes = C.b200s_encoded_nodes(enc.p)
	}
	opts := C.b200s_solve_opts_t{want_masks: 1}
	if rc := C.b200s_solve(c.ctx, C.uint64_t(id), C.b200s_encoded_pods(enc.p), nodes, C.b200s_encoded_templates(enc.p), &opts,
		&res.c); rc != 0 {
		return nil, fmt.Errorf("%w: b200s_solve=%d %s", ErrFallback, int(rc), c.lastError())
	}
	res.decode(enc)
	for p, r := range m.Relaxable { // the CPU scheduler would relax this pod's preferences and try again (scheduler.go:434-465)
		if r && res.PodKind[p] == PlaceUnscheduled {
			return nil, fmt.Errorf("%w: pod %d failed with relaxable preferences", ErrFallback, p)
		}
	}
	// ToNodeClaim's instance-type list (nodeclaimtemplate.go:77-82): the whole OrderByPrice order per NodeClaim, cut at
	// MaxInstanceTypes; a price tie would make the order depend on Go's unstable sort, so such a batch stays on the CPU
	if err := res.priceOrder(c, id); err != nil {
		return nil, err
	}
	return res, nil
}

// SolveBatch runs independent simulations (SimulateScheduling per consolidation candidate, disruption/helpers.go:52-143)
// as one launch: one CTA per simulation.  All models must come from the same catalog generation and key layout.
func (c *Context) SolveBatch(ms []*Model, generation uint64) ([]*Result, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	if c.ctx == nil || len(ms) == 0 {
		return nil, fmt.Errorf("%w: nothing to do", ErrFallback)
	}
	var pin runtime.Pinner
	defer pin.Unpin()
	encs := make([]*encoded, 0, len(ms))
	defer func() {
		for _, e := range encs {
			e.free()
		}
	}()
	n := len(ms)
	pods := make([]*C.b200s_pods_soa_t, n)
	nodes := make([]*C.b200s_nodes_soa_t, n)
	outs := make([]*C.b200s_result_t, n)
	results := make([]*Result, n)
	var id uint64
	for i, m := range ms {
		e, err := encode(m, &pin)
		if err != nil {
			return nil, err
		}
		encs = append(encs, e)
		if i == 0 {
			if id, err = c.catalogFor(e, generation); err != nil {
				return nil, err
			}
		} else if C.b200s_encoded_catalog(e.p).layout_hash != C.b200s_encoded_catalog(encs[0].p).layout_hash {
			return nil, fmt.Errorf("%w: simulation %d uses another key layout (run it on its own)", ErrFallback, i)
		}
		results[i] = newResult(m, e)
		results[i].pin(&pin)
		pods[i] = C.b200s_encoded_pods(e.p)
		if m.NumNodes() > 0 {
			nodes[i] = C.b200s_encoded_nodes(e.p)
		}
		outs[i] = &results[i].c
	}
	pin.Pin(&pods[0])
	pin.Pin(&nodes[0])
	pin.Pin(&outs[0])
	opts := C.b200s_solve_opts_t{want_masks: 1}
	if rc := C.b200s_solve_batch(c.ctx, C.uint64_t(id), C.int32_t(n), (**C.b200s_pods_soa_t)(unsafe.Pointer(&pods[0])),
		(**C.b200s_nodes_soa_t)(unsafe.Pointer(&nodes[0])), C.b200s_encoded_templates(encs[0].p), &opts,
		(**C.b200s_result_t)(unsafe.Pointer(&outs[0]))); rc != 0 {
		return nil, fmt.Errorf("%w: b200s_solve_batch=%d %s", ErrFallback, int(rc), c.lastError())
	}
	for i := range results {
		results[i].decode(encs[i])
	}
	return results, nil
}

// SolveBatchShared runs the simulations of one consolidation pass (SimulateScheduling per candidate,
// disruption/helpers.go:52-143) from ONE model: every node of the cluster and the candidates' pods pending, one candidate
// after the other.  podOff[b] .. podOff[b+1] are simulation b's pods, removed[b] the model nodes it deletes
// (helpers.go:83-97).  One b200s_encode, the node rows cross the bus once; result b's pod indices count from podOff[b]
// and its existing-node rows are rows of the shared node set (Result.NodeOrder maps them to model nodes).
func (c *Context) SolveBatchShared(m *Model, podOff []int32, removed [][]int32, generation uint64) ([]*Result, error) {
	c.mu.Lock()
	defer c.mu.Unlock()
	n := len(podOff) - 1
	if c.ctx == nil || n < 1 || len(removed) != n {
		return nil, fmt.Errorf("%w: nothing to do", ErrFallback)
	}
	var pin runtime.Pinner
	defer pin.Unpin()
	enc, err := encode(m, &pin)
	if err != nil {
		return nil, err
	}
	defer enc.free()
	id, err := c.catalogFor(enc, generation)
	if err != nil {
		return nil, err
	}
	nodes := C.b200s_encoded_nodes(enc.p)
	rowOf := make(map[int32]int32, int(nodes.n_nodes)) // model node -> row in sortExistingNodes order
	for i := 0; i < int(nodes.n_nodes); i++ {
		rowOf[int32(*(*C.int32_t)(unsafeIndex(nodes.node_index, i)))] = int32(i)
	}
	exclOff := make([]int32, n+1)
	exclRows := make([]int32, 0, n)
	for b, nodesOfB := range removed {
		for _, e := range nodesOfB {
			row, ok := rowOf[e]
			if !ok {
				return nil, fmt.Errorf("%w: simulation %d removes a node that is not in the model", ErrFallback, b)
			}
			exclRows = append(exclRows, row)
		}
		exclOff[b+1] = int32(len(exclRows))
	}
	if len(exclRows) == 0 {
		exclRows = append(exclRows, 0) // a valid pointer; exclOff says nothing is read
	}
	outs := make([]*C.b200s_result_t, n)
	results := make([]*Result, n)
	for b := 0; b < n; b++ {
		results[b] = newResultRange(m, enc, int(podOff[b]), int(podOff[b+1]-podOff[b]))
		results[b].pin(&pin)
		outs[b] = &results[b].c
	}
	pin.Pin(&outs[0])
	pin.Pin(&podOff[0])
	pin.Pin(&exclOff[0])
	pin.Pin(&exclRows[0])
	opts := C.b200s_solve_opts_t{want_masks: 1}
	if rc := C.b200s_solve_batch_shared(c.ctx, C.uint64_t(id), C.int32_t(n), C.b200s_encoded_pods(enc.p),
		(*C.int32_t)(unsafe.Pointer(&podOff[0])), nodes, (*C.int32_t)(unsafe.Pointer(&exclOff[0])),
		(*C.int32_t)(unsafe.Pointer(&exclRows[0])), C.b200s_encoded_templates(enc.p), &opts,
		(**C.b200s_result_t)(unsafe.Pointer(&outs[0]))); rc != 0 {
		return nil, fmt.Errorf("%w: b200s_solve_batch_shared=%d %s", ErrFallback, int(rc), c.lastError())
	}
	for _, r := range results {
		r.decode(enc)
	}
	return results, nil
}

// CreateSelect is instance.Provider.Create's vmSize pick (pkg/providers/instance/instance.go:76-95) with the price order
// the scheduler computed instead of Values()[0] (Go map order): host-only, no device (config 1).
func CreateSelect(name string, instanceTypes []string, prices []float64, storageBytes int64) (chosen int, diskGB int32, gpu bool, err error) {
	cname := C.CString(name)
	defer C.free(unsafe.Pointer(cname))
	cts := make([]*C.char, len(instanceTypes))
	for i, s := range instanceTypes {
		cts[i] = C.CString(s)
		defer C.free(unsafe.Pointer(cts[i]))
	}
	var pp *C.double
	if len(prices) == len(instanceTypes) && len(prices) > 0 {
		pp = (*C.double)(unsafe.Pointer(&prices[0]))
	}
	var tp **C.char
	if len(cts) > 0 {
		tp = (**C.char)(unsafe.Pointer(&cts[0]))
	}
	var ch, disk, g C.int32_t
	if rc := C.b200s_create_select(cname, C.int32_t(len(cts)), tp, pp, C.int64_t(storageBytes), &ch, &disk, &g); rc != 0 {
		return -1, 0, false, fmt.Errorf("b200s_create_select=%d", int(rc)) // B200S_CREATE_EBADNAME / ENOTYPE / ENOSTORAGE
	}
	return int(ch), int32(disk), g != 0, nil
}
