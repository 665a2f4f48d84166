package b200sched

/*
#include "b200sched.h"
*/
import "C"

import (
	"fmt"
	"math"
	"runtime"
	"strconv"

	corev1 "k8s.io/api/core/v1"
)

const (
	PlaceUnscheduled = int32(C.B200S_PLACE_UNSCHEDULED)
	PlaceNode        = int32(C.B200S_PLACE_NODE)
	PlaceClaim       = int32(C.B200S_PLACE_CLAIM)
	// MaxInstanceTypes is nodeclaimtemplate.go:41
	MaxInstanceTypes = 600
)

// Claim is one entry of Results.NewNodeClaims, in the reference's order.
type Claim struct {
	Template      int32                             // index into SchedulerState.Templates (the template that opened it)
	Pods          []int32                           // pod indices in Pods append order
	Requests      []int64                           // per resource of Model.Resources (cpu in milli-units)
	Requirements  []corev1.NodeSelectorRequirement   // merged requirements (FinalizeScheduling applied: no hostname)
	InstanceTypes []int32                           // InstanceTypeOptions in OrderByPrice order, cut at MaxInstanceTypes
	Cheapest      int32                             // OrderByPrice(reqs)[0]
	Price         float64
}

// Result is scheduling.Results in index form; the patched Solve turns it into NodeClaim / ExistingNode objects.
type Result struct {
	c  C.b200s_result_t
	m  *Model
	W  int
	K  int
	NB int
	// PodBase / nPods: the pods of Model this result is about (SolveBatchShared: one simulation's range of the shared
	// model; pod indices in Claims[i].Pods, PodKind, ... count from PodBase).  nPods < 0: all pods of the model.
	PodBase int
	nPods   int

	PodKind  []int32 // PlaceUnscheduled / PlaceNode / PlaceClaim
	PodIndex []int32 // existing-node row (sorted order, see NodeOrder) or position in Claims
	PodSeq   []int32
	PodCode  []uint8 // [P*T] per-template attempt code of the pods left in PodErrors (B200S_FIT_*)
	Claims   []Claim
	// NodeOrder[i] = index into SchedulerState.Nodes of existing-node row i (sortExistingNodes, scheduler.go:729-742)
	NodeOrder []int32

	claimTemplate, claimNPods, claimSku, claimNOpt []int32
	claimRequests, claimGt, claimLt                []int64
	claimMask, claimReqMask, claimReqFlags, price  []uint64
}

func newResult(m *Model, enc *encoded) *Result { return newResultRange(m, enc, 0, -1) }

// newResultRange sizes the buffers for the pods [base, base+nPods) of the model (nPods < 0: all of them).
func newResultRange(m *Model, enc *encoded, base, nPods int) *Result {
	cat := C.b200s_encoded_catalog(enc.p)
	tm := C.b200s_encoded_templates(enc.p)
	P, T, R := nPods, int(tm.n_templates), m.NumRes()
	if P < 0 {
		P = m.NumPods()
	}
	W, K, NB := int(cat.n_words), int(cat.n_keys), int(cat.n_bounded)
	n := P // a NodeClaim holds at least one pod
	if n == 0 {
		n = 1
	}
	r := &Result{m: m, W: W, K: K, NB: NB, PodBase: base, nPods: P,
		PodKind: make([]int32, n), PodIndex: make([]int32, n), PodSeq: make([]int32, n), PodCode: make([]uint8, n*max(T, 1)),
		claimTemplate: make([]int32, n), claimNPods: make([]int32, n), claimSku: make([]int32, n), claimNOpt: make([]int32, n),
		claimRequests: make([]int64, n*R), claimGt: make([]int64, n*max(NB, 1)), claimLt: make([]int64, n*max(NB, 1)),
		claimMask: make([]uint64, n*W), claimReqMask: make([]uint64, n*max(K, 1)), claimReqFlags: make([]uint64, n), price: make([]uint64, n)}
	r.c.cap_claims = C.int32_t(n)
	return r
}

// pin points the C result structure at the Go buffers for the duration of the call.
func (r *Result) pin(pin *runtime.Pinner) {
	r.c.pod_place_kind = (*C.int32_t)(ptr(pin, r.PodKind))
	r.c.pod_place_index = (*C.int32_t)(ptr(pin, r.PodIndex))
	r.c.pod_seq = (*C.int32_t)(ptr(pin, r.PodSeq))
	r.c.pod_fit_code = (*C.uint8_t)(ptr(pin, r.PodCode))
	r.c.claim_template = (*C.int32_t)(ptr(pin, r.claimTemplate))
	r.c.claim_npods = (*C.int32_t)(ptr(pin, r.claimNPods))
	r.c.claim_requests = (*C.int64_t)(ptr(pin, r.claimRequests))
	r.c.claim_mask = (*C.uint64_t)(ptr(pin, r.claimMask))
	r.c.claim_req_mask = (*C.uint64_t)(ptr(pin, r.claimReqMask))
	r.c.claim_req_flags = (*C.uint64_t)(ptr(pin, r.claimReqFlags))
	r.c.claim_req_gt = (*C.int64_t)(ptr(pin, r.claimGt))
	r.c.claim_req_lt = (*C.int64_t)(ptr(pin, r.claimLt))
	r.c.claim_sku = (*C.int32_t)(ptr(pin, r.claimSku))
	r.c.claim_price = (*C.uint64_t)(ptr(pin, r.price))
	r.c.claim_n_options = (*C.int32_t)(ptr(pin, r.claimNOpt))
}

// decode turns device key slots / value bits back into label keys / values (b200s_encoded_key_model_id,
// b200s_encoded_value_model_id) and groups the pods by NodeClaim in placement order.
func (r *Result) decode(enc *encoded) {
	n := int(r.c.n_claims)
	P, R := r.nPods, r.m.NumRes()
	tm := C.b200s_encoded_templates(enc.p)
	nodes := C.b200s_encoded_nodes(enc.p)
	r.NodeOrder = make([]int32, int(nodes.n_nodes))
	for i := range r.NodeOrder {
		r.NodeOrder[i] = int32(*(*C.int32_t)(unsafeIndex(nodes.node_index, i)))
	}
	r.Claims = make([]Claim, n)
	for i := 0; i < n; i++ {
		c := &r.Claims[i]
		c.Template = int32(tm.pool_index[r.claimTemplate[i]]) // device template slot -> model pool
		c.Requests = append([]int64(nil), r.claimRequests[i*R:(i+1)*R]...)
		c.Cheapest = r.claimSku[i]
		c.Price = math.Float64frombits(r.price[i])
		c.Pods = make([]int32, 0, r.claimNPods[i])
		flags := r.claimReqFlags[i]
		for k := 0; k < r.K; k++ {
			nib := (flags >> (4 * uint(k))) & 0xF
			if nib&1 == 0 { // not defined
				continue
			}
			key := r.m.Keys[int(C.b200s_encoded_key_model_id(enc.p, C.int(k)))]
			mask := r.claimReqMask[i*r.K+k]
			var vals []string
			for b := 0; b < 64; b++ {
				if mask>>uint(b)&1 == 1 {
					vals = append(vals, r.m.Vals[int(C.b200s_encoded_value_model_id(enc.p, C.int(k), C.int(b)))])
				}
			}
			req := corev1.NodeSelectorRequirement{Key: key}
			switch {
			case nib&2 != 0 && nib&4 != 0: // complement with a Gt bound
				req.Operator, req.Values = corev1.NodeSelectorOpGt, []string{strconv.FormatInt(r.claimGt[i*max(r.NB, 1)+bslot(enc, k)], 10)}
			case nib&2 != 0 && nib&8 != 0:
				req.Operator, req.Values = corev1.NodeSelectorOpLt, []string{strconv.FormatInt(r.claimLt[i*max(r.NB, 1)+bslot(enc, k)], 10)}
			case nib&2 != 0 && len(vals) > 0:
				req.Operator, req.Values = corev1.NodeSelectorOpNotIn, vals
			case nib&2 != 0:
				req.Operator = corev1.NodeSelectorOpExists
			case len(vals) > 0:
				req.Operator, req.Values = corev1.NodeSelectorOpIn, vals
			default:
				req.Operator = corev1.NodeSelectorOpDoesNotExist
			}
			c.Requirements = append(c.Requirements, req)
		}
	}
	// pods of each claim in Pods append order (pod_seq is the global placement order)
	order := make([]int32, P)
	cnt := 0
	bySeq := make([]int32, P)
	for p := 0; p < P; p++ {
		bySeq[p] = -1
	}
	for p := 0; p < P; p++ {
		if r.PodKind[p] != PlaceUnscheduled && r.PodSeq[p] >= 0 {
			bySeq[r.PodSeq[p]] = int32(p)
			cnt++
		}
	}
	_ = order
	for s := 0; s < cnt; s++ {
		p := bySeq[s]
		if p >= 0 && r.PodKind[p] == PlaceClaim {
			c := &r.Claims[r.PodIndex[p]]
			c.Pods = append(c.Pods, p)
		}
	}
}

// priceOrder fills Claims[i].InstanceTypes with ToNodeClaim's list (nodeclaimtemplate.go:77-82): OrderByPrice of the
// claim's options under its requirements, cut at MaxInstanceTypes — b200s_order_by_price.  Must be called with c.mu held.
func (r *Result) priceOrder(c *Context, catalogID uint64) error {
	n := int(r.c.n_claims)
	if n == 0 {
		return nil
	}
	var pin runtime.Pinner
	defer pin.Unpin()
	order, count, tie := make([]int32, n*MaxInstanceTypes), make([]int32, n), make([]uint8, n)
	var gt, lt *C.int64_t
	if r.NB > 0 {
		gt, lt = (*C.int64_t)(ptr(&pin, r.claimGt)), (*C.int64_t)(ptr(&pin, r.claimLt))
	}
	if rc := C.b200s_order_by_price(c.ctx, C.uint64_t(catalogID), C.int32_t(n), (*C.uint64_t)(ptr(&pin, r.claimMask)),
		(*C.uint64_t)(ptr(&pin, r.claimReqMask)), (*C.uint64_t)(ptr(&pin, r.claimReqFlags)), gt, lt, C.int32_t(MaxInstanceTypes),
		(*C.int32_t)(ptr(&pin, order)), (*C.int32_t)(ptr(&pin, count)), (*C.uint8_t)(ptr(&pin, tie))); rc != 0 {
		return fmt.Errorf("%w: b200s_order_by_price=%d %s", ErrFallback, int(rc), c.lastError())
	}
	for i := 0; i < n; i++ {
		if tie[i] != 0 { // equal prices: Go's unstable sort decides the order; keep the vendored OrderByPrice
			return fmt.Errorf("%w: price tie among the instance types of NodeClaim %d", ErrFallback, i)
		}
		r.Claims[i].InstanceTypes = append([]int32(nil), order[i*MaxInstanceTypes:i*MaxInstanceTypes+int(count[i])]...)
	}
	return nil
}

func bslot(enc *encoded, k int) int {
	cat := C.b200s_encoded_catalog(enc.p)
	for b := 0; b < int(cat.n_bounded); b++ {
		if int(cat.bounded_key[b]) == k {
			return b
		}
	}
	return 0
}
