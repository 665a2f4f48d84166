package b200sched

// flatten.go — Karpenter objects -> Model (the Level-1 semantic model of include/b200sched.h).  It lives in the
// scheduling package's import graph (it needs the scheduler's unexported state only through the accessor struct
// SchedulerState below, which the patched scheduler.go fills in).
//
// What is restated here, with the reference lines it follows — each rule has a Python twin in
// gpu-provisioner_b200/b200sched/flatten.py that the CPU test suite runs (tests/test_flatten.py):
//   * NewPodRequirements            vendor/sigs.k8s.io/karpenter/pkg/scheduling/requirements.go:74-110
//   * RequestsForPods / podRequests vendor/sigs.k8s.io/karpenter/pkg/utils/resources/resources.go:30-38,114-144
//                                   (delegated to b200s_pod_requests' rule: sum of containers, max with each init
//                                   container, + overhead, + pods: 1)
//   * NewNodeClaimTemplate          .../provisioning/scheduling/nodeclaimtemplate.go:56-75 (requirements = spec
//                                   requirements + labels incl. the nodepool label)
//   * InstanceType.Allocatable      vendor/sigs.k8s.io/karpenter/pkg/cloudprovider/types.go:181-219 (called, not restated)
//   * ExistingNode                  .../provisioning/scheduling/existingnode.go:41-67 (labels as requirements,
//                                   remaining = Available() - remaining daemon resources)
// Anything outside the exact encoding (pod (anti-)affinity, topology spread, host ports, PVCs, DRA, preferences beyond the
// heaviest node-affinity term, minValues) makes Flatten return ErrFallback: SURVEY.md Appendix A.12.

import (
	"fmt"
	"math"
	"sort"
	"strconv"

	corev1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/api/resource"

	v1 "sigs.k8s.io/karpenter/pkg/apis/v1"
	"sigs.k8s.io/karpenter/pkg/cloudprovider"
	"sigs.k8s.io/karpenter/pkg/scheduling"
	"sigs.k8s.io/karpenter/pkg/utils/resources"
)

// TemplateState is what the scheduler knows about one NodeClaimTemplate (scheduler.go:143-172).
type TemplateState struct {
	NodePoolName   string
	NodePoolWeight int32
	Requirements   []v1.NodeSelectorRequirementWithMinValues // nct.Spec.Requirements
	Labels         map[string]string                         // nct.Labels (includes karpenter.sh/nodepool)
	Taints         []corev1.Taint                            // nct.Spec.Taints (StartupTaints are not scheduling taints)
	InstanceTypes  []*cloudprovider.InstanceType             // GetInstanceTypes(nodePool), BEFORE the scheduler's pre-filter
	DaemonOverhead corev1.ResourceList                       // s.daemonOverhead[nct]
	Remaining      corev1.ResourceList                       // s.remainingResources[pool]; nil = no limits
}

// NodeState is one ExistingNode (existingnode.go:41-67).
type NodeState struct {
	Name        string
	Labels      map[string]string
	Taints      []corev1.Taint
	Remaining   corev1.ResourceList // ExistingNode.remainingResources
	Initialized bool
}

// SchedulerState is everything Solve reads besides the pods.
type SchedulerState struct {
	Templates []TemplateState
	Nodes     []NodeState
}

type flattener struct {
	m      *Model
	keyID  map[string]int32
	valID  map[string]int32 // "key\x00value" -> id for label values; "\x01s" for taint / toleration strings
	resID  map[string]int32
	skuID  map[*cloudprovider.InstanceType]int32
	skuSet map[string]int32
}

func newFlattener() *flattener {
	f := &flattener{m: &Model{KeyZone: -1, KeyCapacityType: -1, KeyInstType: -1, ResCPU: -1, ResMemory: -1},
		keyID: map[string]int32{}, valID: map[string]int32{}, resID: map[string]int32{},
		skuID: map[*cloudprovider.InstanceType]int32{}, skuSet: map[string]int32{}}
	// value id 0 is the empty string of taints / tolerations
	f.m.Vals = append(f.m.Vals, "")
	f.m.ValKey = append(f.m.ValKey, -1)
	f.m.ValInt = append(f.m.ValInt, 0)
	f.m.ValIsInt = append(f.m.ValIsInt, 0)
	f.valID["\x01"] = 0
	return f
}

func (f *flattener) key(k string) int32 {
	if n, ok := v1.NormalizedLabels[k]; ok { // requirement.go:48-50
		k = n
	}
	if id, ok := f.keyID[k]; ok {
		return id
	}
	id := int32(len(f.m.Keys))
	f.keyID[k] = id
	f.m.Keys = append(f.m.Keys, k)
	wk := uint8(0)
	if v1.WellKnownLabels.Has(k) { // labels.go:83-92: the allow-set of Compatible on the new-claim path
		wk = 1
	}
	f.m.KeyWellKnown = append(f.m.KeyWellKnown, wk)
	switch k {
	case corev1.LabelTopologyZone:
		f.m.KeyZone = id
	case v1.CapacityTypeLabelKey:
		f.m.KeyCapacityType = id
	case corev1.LabelInstanceTypeStable:
		f.m.KeyInstType = id
	}
	return id
}

func (f *flattener) value(key int32, s string) int32 {
	k := strconv.Itoa(int(key)) + "\x00" + s
	if id, ok := f.valID[k]; ok {
		return id
	}
	id := int32(len(f.m.Vals))
	f.valID[k] = id
	f.m.Vals = append(f.m.Vals, s)
	f.m.ValKey = append(f.m.ValKey, key)
	n, err := strconv.Atoi(s) // withinIntPtrs: requirement.go:311-327
	if err == nil {
		f.m.ValInt = append(f.m.ValInt, int64(n))
		f.m.ValIsInt = append(f.m.ValIsInt, 1)
	} else {
		f.m.ValInt = append(f.m.ValInt, 0)
		f.m.ValIsInt = append(f.m.ValIsInt, 0)
	}
	return id
}

func (f *flattener) str(s string) int32 { // taint / toleration strings share the value table, key -1
	if s == "" {
		return 0
	}
	k := "\x01" + s
	if id, ok := f.valID[k]; ok {
		return id
	}
	id := int32(len(f.m.Vals))
	f.valID[k] = id
	f.m.Vals = append(f.m.Vals, s)
	f.m.ValKey = append(f.m.ValKey, -1)
	f.m.ValInt = append(f.m.ValInt, 0)
	f.m.ValIsInt = append(f.m.ValIsInt, 0)
	return id
}

func (f *flattener) resource(name corev1.ResourceName) int32 {
	if id, ok := f.resID[string(name)]; ok {
		return id
	}
	id := int32(len(f.m.Resources))
	f.resID[string(name)] = id
	f.m.Resources = append(f.m.Resources, string(name))
	switch name {
	case corev1.ResourceCPU:
		f.m.ResCPU = id
	case corev1.ResourceMemory:
		f.m.ResMemory = id
	}
	return id
}

// quantity is the exact fixed-scale int64 of a resource.Quantity: cpu in milli-units, everything else in base units
// (SURVEY A.6).  A quantity that is not an integer in that unit, or overflows, cannot be encoded.
func quantity(name corev1.ResourceName, q resource.Quantity) (int64, error) {
	if name == corev1.ResourceCPU {
		if q.Cmp(*resource.NewMilliQuantity(q.MilliValue(), q.Format)) != 0 {
			return 0, fmt.Errorf("%w: cpu quantity %s is finer than 1m", ErrFallback, q.String())
		}
		return q.MilliValue(), nil
	}
	v, ok := q.AsInt64()
	if !ok {
		if q.Cmp(*resource.NewQuantity(q.Value(), q.Format)) != 0 {
			return 0, fmt.Errorf("%w: quantity %s of %s is not an integer", ErrFallback, q.String(), name)
		}
		v = q.Value()
	}
	if v > math.MaxInt64/4 || v < -math.MaxInt64/4 {
		return 0, fmt.Errorf("%w: quantity %s of %s could overflow a sum", ErrFallback, q.String(), name)
	}
	return v, nil
}

// req appends one NodeSelectorRequirement row (NewRequirementWithFlexibility, requirement.go:47-87).
func (f *flattener) req(r corev1.NodeSelectorRequirement, minValues *int) error {
	k := f.key(r.Key)
	row := Req{key: C_int32(k), min_values: -1}
	if minValues != nil {
		return fmt.Errorf("%w: minValues on %s", ErrFallback, r.Key)
	}
	switch r.Operator {
	case corev1.NodeSelectorOpIn, corev1.NodeSelectorOpNotIn:
		row.op = opCode(r.Operator)
		row.val_off = C_int32(int32(len(f.m.ReqValues)))
		row.val_cnt = C_int32(int32(len(r.Values)))
		for _, s := range r.Values {
			f.m.ReqValues = append(f.m.ReqValues, f.value(k, s))
		}
	case corev1.NodeSelectorOpExists, corev1.NodeSelectorOpDoesNotExist:
		row.op = opCode(r.Operator)
	case corev1.NodeSelectorOpGt, corev1.NodeSelectorOpLt:
		if len(r.Values) != 1 {
			return fmt.Errorf("%w: %s needs exactly one value", ErrFallback, r.Operator)
		}
		n, err := strconv.Atoi(r.Values[0]) // requirement.go:77,82: an unparsable bound panics there
		if err != nil {
			return fmt.Errorf("%w: %s bound %q", ErrFallback, r.Operator, r.Values[0])
		}
		row.op = opCode(r.Operator)
		row.bound = C_int64(int64(n))
	default:
		return fmt.Errorf("%w: operator %q", ErrFallback, r.Operator)
	}
	f.m.Reqs = append(f.m.Reqs, row)
	return nil
}

// labelReqs appends `key In [value]` for every label (NewLabelRequirements, requirements.go:58-64) in key order.
func (f *flattener) labelReqs(labels map[string]string) error {
	keys := make([]string, 0, len(labels))
	for k := range labels {
		keys = append(keys, k)
	}
	sort.Strings(keys)
	for _, k := range keys {
		if err := f.req(corev1.NodeSelectorRequirement{Key: k, Operator: corev1.NodeSelectorOpIn, Values: []string{labels[k]}}, nil); err != nil {
			return err
		}
	}
	return nil
}

// podReqs is NewPodRequirements (requirements.go:74-110): nodeSelector AND the heaviest preferred node-affinity term
// AND the first required term.  Rows on one key are intersected by the encoder exactly as Requirements.Add does.
func (f *flattener) podReqs(pod *corev1.Pod) error {
	if err := f.labelReqs(pod.Spec.NodeSelector); err != nil {
		return err
	}
	if pod.Spec.Affinity == nil || pod.Spec.Affinity.NodeAffinity == nil {
		return nil
	}
	na := pod.Spec.Affinity.NodeAffinity
	if pref := na.PreferredDuringSchedulingIgnoredDuringExecution; len(pref) > 0 {
		// sort.Slice by weight descending is unstable; up to 12 elements it is an insertion sort, i.e. stable: the
		// first term of maximal weight.  Longer lists with a tie at the top stay on the CPU path.
		best := 0
		for i := 1; i < len(pref); i++ {
			if pref[i].Weight > pref[best].Weight {
				best = i
			}
		}
		if len(pref) > 12 {
			for i := range pref {
				if i != best && pref[i].Weight == pref[best].Weight {
					return fmt.Errorf("%w: tie among more than 12 preferred node-affinity terms", ErrFallback)
				}
			}
		}
		if len(pref[best].Preference.MatchFields) > 0 {
			return fmt.Errorf("%w: matchFields", ErrFallback)
		}
		for _, e := range pref[best].Preference.MatchExpressions {
			if err := f.req(e, nil); err != nil {
				return err
			}
		}
	}
	if req := na.RequiredDuringSchedulingIgnoredDuringExecution; req != nil && len(req.NodeSelectorTerms) > 0 {
		for _, e := range req.NodeSelectorTerms[0].MatchExpressions {
			if err := f.req(e, nil); err != nil {
				return err
			}
		}
	}
	return nil
}

// plainPod reports whether the un-kernelised branches of the scheduler are inert for this pod (SURVEY A.12).
func plainPod(pod *corev1.Pod) error {
	if pod.Spec.Affinity != nil && (pod.Spec.Affinity.PodAffinity != nil || pod.Spec.Affinity.PodAntiAffinity != nil) {
		return fmt.Errorf("%w: pod (anti-)affinity", ErrFallback)
	}
	if len(pod.Spec.TopologySpreadConstraints) > 0 {
		return fmt.Errorf("%w: topology spread", ErrFallback)
	}
	if len(pod.Spec.ResourceClaims) > 0 {
		return fmt.Errorf("%w: resource claims", ErrFallback)
	}
	for _, v := range pod.Spec.Volumes {
		if v.PersistentVolumeClaim != nil || v.Ephemeral != nil {
			return fmt.Errorf("%w: persistent volumes", ErrFallback)
		}
	}
	for _, c := range append(append([]corev1.Container{}, pod.Spec.Containers...), pod.Spec.InitContainers...) {
		for _, p := range c.Ports {
			if p.HostPort != 0 {
				return fmt.Errorf("%w: host ports", ErrFallback)
			}
		}
	}
	return nil
}

// relaxable reports whether Preferences.Relax (preferences.go:38-57) could change this pod's requirements after a failed
// attempt: a preferred node-affinity term, or more than one required term.  The device evaluates NewPodRequirements'
// first answer only, so such a pod must schedule on the first try (Context.Solve checks, else ErrFallback).
func relaxable(pod *corev1.Pod) bool {
	if pod.Spec.Affinity == nil || pod.Spec.Affinity.NodeAffinity == nil {
		return false
	}
	na := pod.Spec.Affinity.NodeAffinity
	if len(na.PreferredDuringSchedulingIgnoredDuringExecution) > 0 {
		return true
	}
	r := na.RequiredDuringSchedulingIgnoredDuringExecution
	return r != nil && len(r.NodeSelectorTerms) > 1
}

func (f *flattener) taints(ts []corev1.Taint) {
	for _, t := range ts {
		f.m.Taints = append(f.m.Taints, Taint{key: C_int32(f.str(t.Key)), value: C_int32(f.str(t.Value)), effect: effectCode(t.Effect)})
	}
}

func (f *flattener) tolerations(ts []corev1.Toleration) {
	for _, t := range ts {
		op := int32(3) // B200S_TOLOP_OTHER
		switch t.Operator {
		case "":
			op = 0
		case corev1.TolerationOpEqual:
			op = 1
		case corev1.TolerationOpExists:
			op = 2
		}
		f.m.Tolerations = append(f.m.Tolerations, Toleration{key: C_int32(f.str(t.Key)), value: C_int32(f.str(t.Value)), op: C_int32(op),
			effect: effectCode(t.Effect)})
	}
}

// resourceRow writes a ResourceList as a dense row over the model's resource axis (missing = 0).
func (f *flattener) resourceRow(dst []int64, rl corev1.ResourceList) error {
	for name, q := range rl {
		v, err := quantity(name, q)
		if err != nil {
			return err
		}
		dst[f.resource(name)] = v
	}
	return nil
}

// Flatten builds the Model of one Solve: the union of every template's instance types is the catalog, a template's own
// list becomes its `instance-type In [...]` requirement.
func Flatten(st *SchedulerState, pods []*corev1.Pod) (*Model, error) {
	f := newFlattener()
	m := f.m
	// ---- resource axis: fixed order first (cpu, memory), then whatever shows up ----
	f.resource(corev1.ResourceCPU)
	f.resource(corev1.ResourceMemory)
	f.resource(corev1.ResourcePods)
	collect := func(rl corev1.ResourceList) {
		names := make([]string, 0, len(rl))
		for n := range rl {
			names = append(names, string(n))
		}
		sort.Strings(names)
		for _, n := range names {
			f.resource(corev1.ResourceName(n))
		}
	}
	var its []*cloudprovider.InstanceType
	for _, t := range st.Templates {
		for _, it := range t.InstanceTypes {
			if _, ok := f.skuSet[it.Name]; !ok {
				f.skuSet[it.Name] = int32(len(its))
				its = append(its, it)
				collect(it.Capacity)
			}
		}
		collect(t.DaemonOverhead)
	}
	for _, p := range pods {
		collect(resources.RequestsForPods(p))
	}
	R := len(m.Resources)
	if R > 8 {
		return nil, fmt.Errorf("%w: %d resource dimensions", ErrFallback, R)
	}
	// ---- instance types ----
	m.SkuReqOff = append(m.SkuReqOff, 0)
	m.SkuOffOff = append(m.SkuOffOff, 0)
	m.OffReqOff = append(m.OffReqOff, 0)
	for _, it := range its {
		m.SkuNames = append(m.SkuNames, it.Name)
		for _, r := range it.Requirements.NodeSelectorRequirements() {
			if err := f.req(r.NodeSelectorRequirement, r.MinValues); err != nil {
				return nil, err
			}
		}
		m.SkuReqOff = append(m.SkuReqOff, int32(len(m.Reqs)))
		alloc, capa := make([]int64, R), make([]int64, R)
		if err := f.resourceRow(alloc, it.Allocatable()); err != nil {
			return nil, err
		}
		if err := f.resourceRow(capa, it.Capacity); err != nil {
			return nil, err
		}
		m.SkuAlloc = append(m.SkuAlloc, alloc...)
		m.SkuCapacity = append(m.SkuCapacity, capa...)
		for _, o := range it.Offerings {
			if o.ReservationCapacity != 0 || o.Requirements.Get(v1.CapacityTypeLabelKey).Has(v1.CapacityTypeReserved) {
				return nil, fmt.Errorf("%w: reserved offerings", ErrFallback)
			}
			for _, r := range o.Requirements.NodeSelectorRequirements() {
				if err := f.req(r.NodeSelectorRequirement, r.MinValues); err != nil {
					return nil, err
				}
			}
			m.OffReqOff = append(m.OffReqOff, int32(len(m.Reqs)))
			m.OffPrice = append(m.OffPrice, o.Price)
			av := uint8(0)
			if o.Available {
				av = 1
			}
			m.OffAvailable = append(m.OffAvailable, av)
		}
		m.SkuOffOff = append(m.SkuOffOff, int32(len(m.OffPrice)))
	}
	// ---- NodePools (templates), unsorted: OrderByWeight is part of the path ----
	m.PoolReqOff = append(m.PoolReqOff, 0)
	m.PoolTaintOff = append(m.PoolTaintOff, 0)
	names := make([]string, len(st.Templates))
	for i, t := range st.Templates {
		names[i] = t.NodePoolName
	}
	rank := rankOf(names)
	for i, t := range st.Templates {
		m.PoolNames = append(m.PoolNames, t.NodePoolName)
		for _, r := range t.Requirements {
			if err := f.req(r.NodeSelectorRequirement, r.MinValues); err != nil {
				return nil, err
			}
		}
		if err := f.labelReqs(t.Labels); err != nil {
			return nil, err
		}
		// the pool's own instance-type list, when it is narrower than the catalog
		if len(t.InstanceTypes) != len(its) {
			vals := make([]string, len(t.InstanceTypes))
			for j, it := range t.InstanceTypes {
				vals[j] = it.Name
			}
			sort.Strings(vals)
			if err := f.req(corev1.NodeSelectorRequirement{Key: corev1.LabelInstanceTypeStable, Operator: corev1.NodeSelectorOpIn, Values: vals}, nil); err != nil {
				return nil, err
			}
		}
		m.PoolReqOff = append(m.PoolReqOff, int32(len(m.Reqs)))
		for _, tt := range t.Taints {
			if tt.Effect == corev1.TaintEffectPreferNoSchedule {
				return nil, fmt.Errorf("%w: PreferNoSchedule taint (preference relaxation)", ErrFallback)
			}
		}
		f.taints(t.Taints)
		m.PoolTaintOff = append(m.PoolTaintOff, int32(len(m.Taints)))
		m.PoolWeight = append(m.PoolWeight, t.NodePoolWeight)
		m.PoolNameRank = append(m.PoolNameRank, rank[i])
		row := make([]int64, R)
		if err := f.resourceRow(row, t.DaemonOverhead); err != nil {
			return nil, err
		}
		m.PoolDaemon = append(m.PoolDaemon, row...)
		has, lim := make([]uint8, R), make([]int64, R)
		for name, q := range t.Remaining {
			v, err := quantity(name, q)
			if err != nil {
				return nil, err
			}
			r := f.resource(name)
			if int(r) >= R {
				return nil, fmt.Errorf("%w: limit on a resource no instance type or pod names", ErrFallback)
			}
			has[r], lim[r] = 1, v
		}
		m.PoolHasLimit = append(m.PoolHasLimit, has...)
		m.PoolLimit = append(m.PoolLimit, lim...)
	}
	// ---- pods ----
	m.PodReqOff = append(m.PodReqOff, int32(len(m.Reqs)))
	m.PodTolOff = append(m.PodTolOff, 0)
	for _, p := range pods {
		if err := plainPod(p); err != nil {
			return nil, err
		}
		if err := f.podReqs(p); err != nil {
			return nil, err
		}
		m.PodReqOff = append(m.PodReqOff, int32(len(m.Reqs)))
		f.tolerations(p.Spec.Tolerations)
		m.PodTolOff = append(m.PodTolOff, int32(len(m.Tolerations)))
		row := make([]int64, R)
		if err := f.resourceRow(row, resources.RequestsForPods(p)); err != nil { // includes pods: 1 (resources.go:36)
			return nil, err
		}
		m.PodRequests = append(m.PodRequests, row...)
		m.PodTS = append(m.PodTS, p.CreationTimestamp.Unix()) // queue.go:95-104 compares at one-second resolution
		hi, lo, err := uidWords(string(p.UID))
		if err != nil {
			return nil, err
		}
		m.PodUIDHi = append(m.PodUIDHi, hi)
		m.PodUIDLo = append(m.PodUIDLo, lo)
		m.Relaxable = append(m.Relaxable, relaxable(p))
	}
	// ---- existing nodes, unsorted: sortExistingNodes is part of the path ----
	m.NodeReqOff = append(m.NodeReqOff, int32(len(m.Reqs)))
	m.NodeTaintOff = append(m.NodeTaintOff, int32(len(m.Taints)))
	nn := make([]string, len(st.Nodes))
	for i, n := range st.Nodes {
		nn[i] = n.Name
	}
	nrank := rankOf(nn)
	for i, n := range st.Nodes {
		if err := f.labelReqs(n.Labels); err != nil {
			return nil, err
		}
		m.NodeReqOff = append(m.NodeReqOff, int32(len(m.Reqs)))
		f.taints(n.Taints)
		m.NodeTaintOff = append(m.NodeTaintOff, int32(len(m.Taints)))
		row := make([]int64, R)
		if err := f.resourceRow(row, n.Remaining); err != nil {
			return nil, err
		}
		m.NodeAvailable = append(m.NodeAvailable, row...)
		ini := uint8(0)
		if n.Initialized {
			ini = 1
		}
		m.NodeInitialized = append(m.NodeInitialized, ini)
		m.NodeNameRank = append(m.NodeNameRank, nrank[i])
	}
	if len(m.Resources) != R {
		return nil, fmt.Errorf("%w: a resource appeared after the axis was fixed", ErrFallback)
	}
	return m, nil
}

// rankOf returns each string's rank in ascending order (ties cannot happen: names are unique).
func rankOf(names []string) []int32 {
	idx := make([]int, len(names))
	for i := range idx {
		idx[i] = i
	}
	sort.Slice(idx, func(a, b int) bool { return names[idx[a]] < names[idx[b]] })
	out := make([]int32, len(names))
	for r, i := range idx {
		out[i] = int32(r)
	}
	return out
}

// uidWords turns a UUID string into the 128-bit big-endian integer whose order equals the string order of canonical
// lower-case UUIDs (queue.go:105-107 compares UIDs as strings).
func uidWords(uid string) (hi, lo uint64, err error) {
	var n int
	for i := 0; i < len(uid); i++ {
		c := uid[i]
		var d uint64
		switch {
		case c == '-':
			continue
		case c >= '0' && c <= '9':
			d = uint64(c - '0')
		case c >= 'a' && c <= 'f':
			d = uint64(c-'a') + 10
		default:
			return 0, 0, fmt.Errorf("%w: pod UID %q is not a canonical UUID", ErrFallback, uid)
		}
		if n < 16 {
			hi = hi<<4 | d
		} else {
			lo = lo<<4 | d
		}
		n++
	}
	if n != 32 {
		return 0, 0, fmt.Errorf("%w: pod UID %q is not a canonical UUID", ErrFallback, uid)
	}
	return hi, lo, nil
}

// keep the scheduling import used when requirements are passed through unchanged by callers of this package
var _ = scheduling.NewRequirements
