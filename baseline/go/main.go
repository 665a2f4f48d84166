// Off-box cross-check of the oracle: runs the UNMODIFIED vendored Karpenter v1.7.0 scheduler on a model written by
// tools/dump_model.py and prints its result as JSON for tools/compare_go_result.py.
//
// NOT EXECUTED IN THE AUTHORING IMAGE (no Go toolchain there; SURVEY.md section 8c).  It is written against the
// vendored sources under vendor/sigs.k8s.io/karpenter/pkg and needs a Go >= 1.24 toolchain:
//
//	cp -r baseline/go <gpu-provisioner checkout>/hack/b200-baseline
//	cd <gpu-provisioner checkout> && go run -mod=vendor ./hack/b200-baseline model.json > go_result.json
//	python tools/compare_go_result.py model.json go_result.json        # in this repository
//
// Call shape: scheduling.NewTopology + scheduling.NewScheduler + (*Scheduler).Solve, the way
// Provisioner.NewScheduler / Schedule do (pkg/controllers/provisioning/provisioner.go:236-345), with NodePools in
// OrderByWeight order, existing nodes as state.StateNode fixtures, one daemonset pod per pool with overhead, no volumes.  It also times Solve, which is the true
// "Go on host cores" baseline SURVEY.md section 8d (4) asks for.
package main

import (
	"context"
	"encoding/json"
	"fmt"
	"os"
	"sort"
	"time"

	"github.com/awslabs/operatorpkg/status"
	corev1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/api/resource"
	metav1 "k8s.io/apimachinery/pkg/apis/meta/v1"
	"k8s.io/apimachinery/pkg/types"
	"k8s.io/utils/clock"
	"sigs.k8s.io/controller-runtime/pkg/client/fake"

	v1 "sigs.k8s.io/karpenter/pkg/apis/v1"
	"sigs.k8s.io/karpenter/pkg/cloudprovider"
	"sigs.k8s.io/karpenter/pkg/controllers/provisioning/scheduling"
	"sigs.k8s.io/karpenter/pkg/controllers/state"
	"sigs.k8s.io/karpenter/pkg/operator/injection"
	"sigs.k8s.io/karpenter/pkg/operator/options"
	karpsched "sigs.k8s.io/karpenter/pkg/scheduling"
	"sigs.k8s.io/karpenter/pkg/test"
	nodepoolutils "sigs.k8s.io/karpenter/pkg/utils/nodepool"
)

type reqJSON struct {
	Key      string   `json:"key"`
	Operator string   `json:"operator"`
	Values   []string `json:"values"`
}

type taintJSON struct {
	Key    string `json:"key"`
	Value  string `json:"value"`
	Effect string `json:"effect"`
}

type tolerationJSON struct {
	Key      string `json:"key"`
	Operator string `json:"operator"`
	Value    string `json:"value"`
	Effect   string `json:"effect"`
}

type offeringJSON struct {
	Zone         string  `json:"zone"`
	CapacityType string  `json:"capacity_type"`
	Price        float64 `json:"price"`
	Available    bool    `json:"available"`
}

type instanceTypeJSON struct {
	Name         string            `json:"name"`
	Requirements []reqJSON         `json:"requirements"`
	Capacity     map[string]string `json:"capacity"`
	Allocatable  map[string]string `json:"allocatable"`
	Offerings    []offeringJSON    `json:"offerings"`
}

type nodePoolJSON struct {
	Name           string            `json:"name"`
	Weight         int32             `json:"weight"`
	Requirements   []reqJSON         `json:"requirements"`
	Taints         []taintJSON       `json:"taints"`
	Limits         map[string]string `json:"limits"`          // remaining resources of the pool at the start of the Solve
	DaemonOverhead map[string]string `json:"daemon_overhead"` // what getDaemonOverhead must come to for this pool
}

// an existing node: becomes a state.StateNode (Node only when initialized; Node + NodeClaim, registered but not
// initialized, otherwise — StateNode.Initialized, statenode.go:342-349)
type nodeJSON struct {
	Name        string            `json:"name"`
	Labels      map[string]string `json:"labels"`
	Available   map[string]string `json:"available"`
	Taints      []taintJSON       `json:"taints"`
	Initialized bool              `json:"initialized"`
}

type podJSON struct {
	Index             int               `json:"index"`
	UID               string            `json:"uid"`
	CreationTimestamp int64             `json:"creation_timestamp"`
	Requests          map[string]string `json:"requests"`
	Requirements      []reqJSON         `json:"requirements"`
	Tolerations       []tolerationJSON  `json:"tolerations"`
}

type modelJSON struct {
	Resources      []string           `json:"resources"`
	WellKnownExtra []string           `json:"wellknown_extra"`
	InstanceTypes  []instanceTypeJSON `json:"instance_types"`
	NodePools      []nodePoolJSON     `json:"node_pools"`
	Pods           []podJSON          `json:"pods"`
	Nodes          []nodeJSON         `json:"nodes"`
}

type claimJSON struct {
	NodePool      string               `json:"node_pool"`
	Pods          []int                `json:"pods"`
	InstanceTypes []string             `json:"instance_types"`
	Cheapest      string               `json:"cheapest"`
	Requests      map[string]string    `json:"requests"`
	Requirements  map[string]reqOutput `json:"requirements"`
}

type reqOutput struct {
	Operator string   `json:"operator"`
	Values   []string `json:"values"`
}

type resultJSON struct {
	SolveSeconds  float64          `json:"solve_seconds"`
	NewNodeClaims []claimJSON      `json:"new_node_claims"`
	Unschedulable []int            `json:"unschedulable"`
	ExistingNodes map[string][]int `json:"existing_nodes"` // Results.ExistingNodes: node name -> pods it took
}

// the scheduler only stores the provider on state.Cluster; nothing on the Solve path calls it
type stubProvider struct{}

func (stubProvider) Create(context.Context, *v1.NodeClaim) (*v1.NodeClaim, error) {
	return nil, fmt.Errorf("not implemented")
}
func (stubProvider) Delete(context.Context, *v1.NodeClaim) error { return fmt.Errorf("not implemented") }
func (stubProvider) Get(context.Context, string) (*v1.NodeClaim, error) {
	return nil, fmt.Errorf("not implemented")
}
func (stubProvider) List(context.Context) ([]*v1.NodeClaim, error) { return nil, nil }
func (stubProvider) GetInstanceTypes(context.Context, *v1.NodePool) ([]*cloudprovider.InstanceType, error) {
	return nil, nil
}
func (stubProvider) IsDrifted(context.Context, *v1.NodeClaim) (cloudprovider.DriftReason, error) {
	return "", nil
}
func (stubProvider) RepairPolicies() []cloudprovider.RepairPolicy { return nil }
func (stubProvider) Name() string                                 { return "b200-baseline" }
func (stubProvider) GetSupportedNodeClasses() []status.Object     { return nil }

func resourceList(m map[string]string, skip string) corev1.ResourceList {
	out := corev1.ResourceList{}
	for k, v := range m {
		if k == skip {
			continue
		}
		out[corev1.ResourceName(k)] = resource.MustParse(v)
	}
	return out
}

func nodeSelectorReqs(rs []reqJSON, skipKey string) []corev1.NodeSelectorRequirement {
	out := []corev1.NodeSelectorRequirement{}
	for _, r := range rs {
		if r.Key == skipKey {
			continue
		}
		out = append(out, corev1.NodeSelectorRequirement{Key: r.Key, Operator: corev1.NodeSelectorOperator(r.Operator), Values: r.Values})
	}
	return out
}

func main() {
	if len(os.Args) != 2 {
		fmt.Fprintln(os.Stderr, "usage: b200-baseline model.json")
		os.Exit(2)
	}
	raw, err := os.ReadFile(os.Args[1])
	if err != nil {
		panic(err)
	}
	var m modelJSON
	if err := json.Unmarshal(raw, &m); err != nil {
		panic(err)
	}
	ctx := options.ToContext(injection.WithControllerName(context.Background(), "provisioner"), test.Options())
	// provider labels the cloud provider registers as well-known (AllowUndefinedWellKnownLabels reads this set)
	v1.WellKnownLabels.Insert(m.WellKnownExtra...)

	// ---- instance types: Allocatable() = Capacity - Overhead.Total() (cloudprovider/types.go:181-214), so the whole
	// difference goes into KubeReserved and Allocatable() returns the dumped allocatable exactly
	instanceTypes := []*cloudprovider.InstanceType{}
	for _, it := range m.InstanceTypes {
		reqs := karpsched.NewRequirements()
		for _, r := range it.Requirements {
			reqs.Add(karpsched.NewRequirement(r.Key, corev1.NodeSelectorOperator(r.Operator), r.Values...))
		}
		offerings := cloudprovider.Offerings{}
		for _, o := range it.Offerings {
			offerings = append(offerings, &cloudprovider.Offering{
				Requirements: karpsched.NewRequirements(
					karpsched.NewRequirement(corev1.LabelTopologyZone, corev1.NodeSelectorOpIn, o.Zone),
					karpsched.NewRequirement(v1.CapacityTypeLabelKey, corev1.NodeSelectorOpIn, o.CapacityType)),
				Price: o.Price, Available: o.Available})
		}
		capacity := resourceList(it.Capacity, "")
		reserved := corev1.ResourceList{}
		for name, c := range capacity {
			a := resource.MustParse(it.Allocatable[string(name)])
			d := c.DeepCopy()
			d.Sub(a)
			reserved[name] = d
		}
		instanceTypes = append(instanceTypes, &cloudprovider.InstanceType{
			Name: it.Name, Requirements: reqs, Offerings: offerings, Capacity: capacity,
			Overhead: &cloudprovider.InstanceTypeOverhead{KubeReserved: reserved}})
	}

	// ---- node pools, in OrderByWeight order (provisioner.go:245)
	nodePools := []*v1.NodePool{}
	byPool := map[string][]*cloudprovider.InstanceType{}
	for _, np := range m.NodePools {
		weight := np.Weight
		reqs := []v1.NodeSelectorRequirementWithMinValues{}
		// NewNodeClaimTemplate adds karpenter.sh/nodepool = name itself (nodeclaimtemplate.go:67-72)
		for _, r := range nodeSelectorReqs(np.Requirements, v1.NodePoolLabelKey) {
			reqs = append(reqs, v1.NodeSelectorRequirementWithMinValues{NodeSelectorRequirement: r})
		}
		taints := []corev1.Taint{}
		for _, t := range np.Taints {
			taints = append(taints, corev1.Taint{Key: t.Key, Value: t.Value, Effect: corev1.TaintEffect(t.Effect)})
		}
		pool := &v1.NodePool{ObjectMeta: metav1.ObjectMeta{Name: np.Name, UID: types.UID("pool-" + np.Name)}}
		pool.Spec.Weight = &weight
		pool.Spec.Template.Spec.Requirements = reqs
		pool.Spec.Template.Spec.Taints = taints
		pool.Spec.Template.Spec.NodeClassRef = &v1.NodeClassReference{Group: "karpenter.test.sh", Kind: "TestNodeClass", Name: "default"}
		if len(np.Limits) > 0 { // NewScheduler: remainingResources[pool] = Spec.Limits (scheduler.go:170-172)
			pool.Spec.Limits = v1.Limits(resourceList(np.Limits, ""))
		}
		nodePools = append(nodePools, pool)
		byPool[np.Name] = instanceTypes
	}
	nodepoolutils.OrderByWeight(nodePools)

	// ---- daemon overhead (getDaemonOverhead, scheduler.go:773-783): one daemonset pod per pool that asks for it, bound
	// to the pool by a node selector on karpenter.sh/nodepool and tolerating everything.  RequestsForPods counts the pod
	// itself as pods: 1, so an overhead is expressible only if its pods entry is 1 (tools/dump_model.py checks).
	daemonSetPods := []*corev1.Pod{}
	for _, np := range m.NodePools {
		reqs := resourceList(np.DaemonOverhead, "pods")
		nonZero := false
		for _, q := range reqs {
			nonZero = nonZero || !q.IsZero()
		}
		if !nonZero && np.DaemonOverhead["pods"] != "1" {
			continue
		}
		daemonSetPods = append(daemonSetPods, &corev1.Pod{
			ObjectMeta: metav1.ObjectMeta{Name: "daemon-" + np.Name, Namespace: "kube-system", UID: types.UID("daemon-" + np.Name)},
			Spec: corev1.PodSpec{
				NodeSelector: map[string]string{v1.NodePoolLabelKey: np.Name},
				Tolerations:  []corev1.Toleration{{Operator: corev1.TolerationOpExists}},
				Containers:   []corev1.Container{{Name: "d", Resources: corev1.ResourceRequirements{Requests: reqs}}}}})
	}

	// ---- pods: requests on one container (RequestsForPods adds pods: 1 itself, resources.go:36), the requirement
	// set as the one required node-affinity term, tolerations as given
	pods := []*corev1.Pod{}
	indexOf := map[types.UID]int{}
	for _, p := range m.Pods {
		pod := &corev1.Pod{ObjectMeta: metav1.ObjectMeta{
			Name: fmt.Sprintf("pod-%07d", p.Index), Namespace: "default", UID: types.UID(p.UID),
			CreationTimestamp: metav1.NewTime(time.Unix(p.CreationTimestamp, 0))}}
		pod.Spec.Containers = []corev1.Container{{Name: "c", Resources: corev1.ResourceRequirements{Requests: resourceList(p.Requests, "pods")}}}
		if exprs := nodeSelectorReqs(p.Requirements, ""); len(exprs) > 0 {
			pod.Spec.Affinity = &corev1.Affinity{NodeAffinity: &corev1.NodeAffinity{
				RequiredDuringSchedulingIgnoredDuringExecution: &corev1.NodeSelector{
					NodeSelectorTerms: []corev1.NodeSelectorTerm{{MatchExpressions: exprs}}}}}
		}
		for _, t := range p.Tolerations {
			pod.Spec.Tolerations = append(pod.Spec.Tolerations, corev1.Toleration{
				Key: t.Key, Operator: corev1.TolerationOperator(t.Operator), Value: t.Value, Effect: corev1.TaintEffect(t.Effect)})
		}
		pods = append(pods, pod)
		indexOf[pod.UID] = p.Index
	}

	// ---- existing nodes as state.StateNode fixtures.  Available() = Allocatable - PodRequests (statenode.go:386-388)
	// and no pod is tracked on the fixture, so Allocatable carries the dumped remaining resources; Capacity is the same
	// (it is read only by updateRemainingResources, for nodes that carry a limited pool's label).  An initialized node
	// is an unmanaged one (no NodeClaim: always Registered and Initialized); an uninitialized node gets an empty
	// NodeClaim and the registered label only.
	stateNodes := []*state.StateNode{}
	for _, n := range m.Nodes {
		labels := map[string]string{}
		for k, v := range n.Labels {
			labels[k] = v
		}
		taints := []corev1.Taint{}
		for _, t := range n.Taints {
			taints = append(taints, corev1.Taint{Key: t.Key, Value: t.Value, Effect: corev1.TaintEffect(t.Effect)})
		}
		sn := state.NewNode()
		if !n.Initialized {
			labels[v1.NodeRegisteredLabelKey] = "true"
			sn.NodeClaim = &v1.NodeClaim{ObjectMeta: metav1.ObjectMeta{Name: n.Name}}
		}
		avail := resourceList(n.Available, "")
		sn.Node = &corev1.Node{
			ObjectMeta: metav1.ObjectMeta{Name: n.Name, Labels: labels},
			Spec:       corev1.NodeSpec{Taints: taints, ProviderID: "b200://" + n.Name},
			Status:     corev1.NodeStatus{Allocatable: avail, Capacity: avail}}
		stateNodes = append(stateNodes, sn)
	}

	kubeClient := fake.NewClientBuilder().Build()
	clk := clock.RealClock{}
	cluster := state.NewCluster(clk, kubeClient, stubProvider{})
	topology, err := scheduling.NewTopology(ctx, kubeClient, cluster, stateNodes, nodePools, byPool, pods)
	if err != nil {
		panic(err)
	}
	scheduler := scheduling.NewScheduler(ctx, kubeClient, nodePools, cluster, stateNodes, topology, byPool, daemonSetPods, test.NewEventRecorder(), clk)
	start := time.Now()
	results, err := scheduler.Solve(ctx, pods)
	elapsed := time.Since(start)
	if err != nil {
		panic(err)
	}

	out := resultJSON{SolveSeconds: elapsed.Seconds(), NewNodeClaims: []claimJSON{}, Unschedulable: []int{}, ExistingNodes: map[string][]int{}}
	for _, en := range results.ExistingNodes {
		for _, p := range en.Pods {
			out.ExistingNodes[en.Name()] = append(out.ExistingNodes[en.Name()], indexOf[p.UID])
		}
		sort.Ints(out.ExistingNodes[en.Name()])
	}
	for _, nc := range results.NewNodeClaims {
		c := claimJSON{NodePool: nc.NodePoolName, Pods: []int{}, InstanceTypes: []string{}, Requests: map[string]string{},
			Requirements: map[string]reqOutput{}}
		for _, p := range nc.Pods {
			c.Pods = append(c.Pods, indexOf[p.UID])
		}
		sort.Ints(c.Pods)
		for _, it := range nc.InstanceTypeOptions {
			c.InstanceTypes = append(c.InstanceTypes, it.Name)
		}
		sort.Strings(c.InstanceTypes)
		if ordered := nc.InstanceTypeOptions.OrderByPrice(nc.Requirements); len(ordered) > 0 { // after the names were copied: it sorts in place
			c.Cheapest = ordered[0].Name
		}
		for name, q := range nc.Spec.Resources.Requests {
			if name == corev1.ResourceCPU {
				c.Requests[string(name)] = fmt.Sprintf("%dm", q.MilliValue())
			} else {
				c.Requests[string(name)] = fmt.Sprintf("%d", q.Value())
			}
		}
		for _, r := range nc.Requirements.NodeSelectorRequirements() {
			vals := append([]string{}, r.Values...)
			sort.Strings(vals)
			c.Requirements[r.Key] = reqOutput{Operator: string(r.Operator), Values: vals}
		}
		out.NewNodeClaims = append(out.NewNodeClaims, c)
	}
	for p := range results.PodErrors {
		out.Unschedulable = append(out.Unschedulable, indexOf[p.UID])
	}
	sort.Ints(out.Unschedulable)
	enc := json.NewEncoder(os.Stdout)
	if err := enc.Encode(out); err != nil {
		panic(err)
	}
}
