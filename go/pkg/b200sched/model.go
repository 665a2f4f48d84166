package b200sched

/*
#include "b200sched.h"
*/
import "C"

import (
	"runtime"
	"unsafe"
)

// Req is one b200s_req_t: key op values|bound.
type Req = C.b200s_req_t

// Taint / Toleration are b200s_taint_t / b200s_toleration_t with interned strings.
type Taint = C.b200s_taint_t
type Toleration = C.b200s_toleration_t

// Model is b200s_model_t with Go-owned backing arrays (the Level-1 semantic model of include/b200sched.h): what flatten.go
// turns Karpenter objects into.  Strings are interned: Keys[i] / Vals[i] / Resources[i] / SkuNames[i] are the strings
// behind the integer ids the C structure carries, kept for results.go's decode.
type Model struct {
	Keys      []string
	Vals      []string // interned label values AND taint / toleration strings; Vals[0] == ""
	Resources []string
	SkuNames  []string
	PoolNames []string

	KeyWellKnown                           []uint8
	KeyZone, KeyCapacityType, KeyInstType  int32
	ValKey                                 []int32
	ValInt                                 []int64
	ValIsInt                               []uint8
	ResCPU, ResMemory                      int32
	Reqs                                   []Req
	ReqValues                              []int32
	Taints                                 []Taint
	Tolerations                            []Toleration
	SkuReqOff                              []int32
	SkuAlloc, SkuCapacity                  []int64
	SkuOffOff                              []int32
	OffReqOff                              []int32
	OffPrice                               []float64
	OffAvailable                           []uint8
	PoolReqOff, PoolTaintOff, PoolWeight   []int32
	PoolNameRank                           []int32
	PoolDaemon                             []int64
	PoolHasLimit                           []uint8
	PoolLimit                              []int64
	PodReqOff, PodTolOff                   []int32
	PodRequests, PodTS                     []int64
	PodUIDHi, PodUIDLo                     []uint64
	NodeReqOff, NodeTaintOff, NodeNameRank []int32
	NodeAvailable                          []int64
	NodeInitialized                        []uint8

	// Go-side only: pods whose requirements Preferences.Relax could loosen after a failed attempt (flatten.go relaxable)
	Relaxable []bool
}

func (m *Model) NumPods() int  { return len(m.PodTS) }
func (m *Model) NumSkus() int  { return len(m.SkuNames) }
func (m *Model) NumPools() int { return len(m.PoolWeight) }
func (m *Model) NumNodes() int { return len(m.NodeInitialized) }
func (m *Model) NumRes() int   { return len(m.Resources) }

// ptr returns the address of a slice's first element (nil for an empty slice) and pins it for the call.
func ptr[T any](pin *runtime.Pinner, s []T) unsafe.Pointer {
	if len(s) == 0 {
		return nil
	}
	pin.Pin(&s[0])
	return unsafe.Pointer(&s[0])
}

// toC fills a b200s_model_t whose pointers refer to the model's slices; pin keeps them in place until Unpin.
func (m *Model) toC(pin *runtime.Pinner) C.b200s_model_t {
	var c C.b200s_model_t
	c.n_keys = C.int32_t(len(m.Keys))
	c.key_wellknown = (*C.uint8_t)(ptr(pin, m.KeyWellKnown))
	c.key_zone, c.key_capacity_type, c.key_instance_type = C.int32_t(m.KeyZone), C.int32_t(m.KeyCapacityType), C.int32_t(m.KeyInstType)
	c.n_vals = C.int32_t(len(m.ValKey))
	c.val_key = (*C.int32_t)(ptr(pin, m.ValKey))
	c.val_int = (*C.int64_t)(ptr(pin, m.ValInt))
	c.val_isint = (*C.uint8_t)(ptr(pin, m.ValIsInt))
	c.n_res = C.int32_t(len(m.Resources))
	c.res_cpu, c.res_memory = C.int32_t(m.ResCPU), C.int32_t(m.ResMemory)
	c.n_reqs = C.int32_t(len(m.Reqs))
	c.reqs = (*C.b200s_req_t)(ptr(pin, m.Reqs))
	c.n_req_values = C.int32_t(len(m.ReqValues))
	c.req_values = (*C.int32_t)(ptr(pin, m.ReqValues))
	c.n_taints = C.int32_t(len(m.Taints))
	c.taints = (*C.b200s_taint_t)(ptr(pin, m.Taints))
	c.n_tolerations = C.int32_t(len(m.Tolerations))
	c.tolerations = (*C.b200s_toleration_t)(ptr(pin, m.Tolerations))
	c.n_skus = C.int32_t(len(m.SkuNames))
	c.sku_req_off = (*C.int32_t)(ptr(pin, m.SkuReqOff))
	c.sku_alloc = (*C.int64_t)(ptr(pin, m.SkuAlloc))
	c.sku_capacity = (*C.int64_t)(ptr(pin, m.SkuCapacity))
	c.sku_off_off = (*C.int32_t)(ptr(pin, m.SkuOffOff))
	c.n_offerings = C.int32_t(len(m.OffPrice))
	c.off_req_off = (*C.int32_t)(ptr(pin, m.OffReqOff))
	c.off_price = (*C.double)(ptr(pin, m.OffPrice))
	c.off_available = (*C.uint8_t)(ptr(pin, m.OffAvailable))
	c.n_pools = C.int32_t(len(m.PoolWeight))
	c.pool_req_off = (*C.int32_t)(ptr(pin, m.PoolReqOff))
	c.pool_taint_off = (*C.int32_t)(ptr(pin, m.PoolTaintOff))
	c.pool_weight = (*C.int32_t)(ptr(pin, m.PoolWeight))
	c.pool_name_rank = (*C.int32_t)(ptr(pin, m.PoolNameRank))
	c.pool_daemon = (*C.int64_t)(ptr(pin, m.PoolDaemon))
	c.pool_has_limit = (*C.uint8_t)(ptr(pin, m.PoolHasLimit))
	c.pool_limit = (*C.int64_t)(ptr(pin, m.PoolLimit))
	c.n_pods = C.int32_t(len(m.PodTS))
	c.pod_req_off = (*C.int32_t)(ptr(pin, m.PodReqOff))
	c.pod_tol_off = (*C.int32_t)(ptr(pin, m.PodTolOff))
	c.pod_requests = (*C.int64_t)(ptr(pin, m.PodRequests))
	c.pod_ts = (*C.int64_t)(ptr(pin, m.PodTS))
	c.pod_uid_hi = (*C.uint64_t)(ptr(pin, m.PodUIDHi))
	c.pod_uid_lo = (*C.uint64_t)(ptr(pin, m.PodUIDLo))
	c.n_nodes = C.int32_t(len(m.NodeInitialized))
	c.node_req_off = (*C.int32_t)(ptr(pin, m.NodeReqOff))
	c.node_taint_off = (*C.int32_t)(ptr(pin, m.NodeTaintOff))
	c.node_available = (*C.int64_t)(ptr(pin, m.NodeAvailable))
	c.node_initialized = (*C.uint8_t)(ptr(pin, m.NodeInitialized))
	c.node_name_rank = (*C.int32_t)(ptr(pin, m.NodeNameRank))
	return c
}
