// Package instancetype is the catalog behind CloudProvider.GetInstanceTypes (SURVEY.md N3): the reference returns an empty
// slice (pkg/cloudprovider/cloudprovider.go:99-101), so the scheduler has nothing to pack onto.  This provider serves a
// static table of Azure GPU SKUs — the names of the reference's own examples/v1-nodeclaim-gpu.yaml:22-48 — as
// cloudprovider.InstanceType rows: requirements for every well-known label (types.go:108-110), one offering per
// (zone, capacity type), capacity and the AKS overhead that Allocatable() subtracts (types.go:181-219).
//
// NOT COMPILED IN THE AUTHORING IMAGE.  gpu-provisioner_b200/b200sched/azure_catalog.py builds the same rows from the same
// JSON for the Python-side tests (tests/test_flatten.py::test_azure_catalog_*).
package instancetype

import (
	_ "embed"
	"encoding/json"
	"fmt"
	"strconv"
	"sync"

	corev1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/api/resource"

	v1 "sigs.k8s.io/karpenter/pkg/apis/v1"
	"sigs.k8s.io/karpenter/pkg/cloudprovider"
	"sigs.k8s.io/karpenter/pkg/scheduling"
)

//go:embed azure_gpu_skus.json
var skuJSON []byte

// Label keys the provider registers as well-known next to Karpenter's own (the Azure provider does the same for its
// sku-* labels): pods may select on them without the NodePool defining them.
const (
	LabelSKUFamily   = "karpenter.azure.com/sku-family"
	LabelSKUCPU      = "karpenter.azure.com/sku-cpu"
	LabelSKUGPUCount = "karpenter.azure.com/sku-gpu-count"
	LabelSKUGPUName  = "karpenter.azure.com/sku-gpu-name"
	SpotDiscount     = 0.3
)

func init() {
	v1.WellKnownLabels.Insert(LabelSKUFamily, LabelSKUCPU, LabelSKUGPUCount, LabelSKUGPUName)
}

type sku struct {
	Name          string  `json:"name"`
	VCPU          int64   `json:"vcpu"`
	MemoryGiB     int64   `json:"memory_gib"`
	GPU           int64   `json:"gpu"`
	GPUName       string  `json:"gpu_name"`
	PriceOnDemand float64 `json:"price_on_demand"`
}

type table struct {
	Zones    []string `json:"zones"`
	SKUs     []sku    `json:"skus"`
	Overhead struct {
		KubeReservedCPUMilli  int64 `json:"kube_reserved_cpu_milli"`
		KubeReservedMemoryMiB int64 `json:"kube_reserved_memory_mib"`
		EvictionMemoryMiB     int64 `json:"eviction_memory_mib"`
		Pods                  int64 `json:"pods"`
		EphemeralGiB          int64 `json:"ephemeral_gib"`
	} `json:"overhead"`
}

// Provider serves the table; Generation changes whenever the served list does (b200sched.Context keys its uploaded
// catalogs by it).
type Provider struct {
	once       sync.Once
	region     string
	types      []*cloudprovider.InstanceType
	err        error
	Generation uint64
}

func NewProvider(region string) *Provider { return &Provider{region: region, Generation: 1} }

// List returns the instance types a NodePool may use.  The list is immutable and shared between Solves.
func (p *Provider) List() ([]*cloudprovider.InstanceType, error) {
	p.once.Do(func() { p.types, p.err = build(p.region) })
	return p.types, p.err
}

func family(name string) string { // Standard_NC6s_v3 -> NC
	s := name[len("Standard_"):]
	i := 0
	for i < len(s) && (s[i] < '0' || s[i] > '9') {
		i++
	}
	return s[:i]
}

func build(region string) ([]*cloudprovider.InstanceType, error) {
	var t table
	if err := json.Unmarshal(skuJSON, &t); err != nil {
		return nil, fmt.Errorf("parsing the SKU table, %w", err)
	}
	var out []*cloudprovider.InstanceType
	for _, s := range t.SKUs {
		zones := make([]string, len(t.Zones))
		for i, z := range t.Zones {
			zones[i] = region + "-" + z
		}
		reqs := scheduling.NewRequirements(
			scheduling.NewRequirement(corev1.LabelInstanceTypeStable, corev1.NodeSelectorOpIn, s.Name),
			scheduling.NewRequirement(corev1.LabelArchStable, corev1.NodeSelectorOpIn, "amd64"),
			scheduling.NewRequirement(corev1.LabelOSStable, corev1.NodeSelectorOpIn, string(corev1.Linux)),
			scheduling.NewRequirement(corev1.LabelTopologyZone, corev1.NodeSelectorOpIn, zones...),
			scheduling.NewRequirement(corev1.LabelTopologyRegion, corev1.NodeSelectorOpIn, region),
			scheduling.NewRequirement(v1.CapacityTypeLabelKey, corev1.NodeSelectorOpIn, v1.CapacityTypeOnDemand, v1.CapacityTypeSpot),
			scheduling.NewRequirement(corev1.LabelWindowsBuild, corev1.NodeSelectorOpDoesNotExist),
			scheduling.NewRequirement(LabelSKUFamily, corev1.NodeSelectorOpIn, family(s.Name)),
			scheduling.NewRequirement(LabelSKUCPU, corev1.NodeSelectorOpIn, strconv.FormatInt(s.VCPU, 10)),
			scheduling.NewRequirement(LabelSKUGPUCount, corev1.NodeSelectorOpIn, strconv.FormatInt(s.GPU, 10)),
			scheduling.NewRequirement(LabelSKUGPUName, corev1.NodeSelectorOpIn, s.GPUName),
		)
		var offs cloudprovider.Offerings
		for zi, z := range zones {
			for _, ct := range []string{v1.CapacityTypeOnDemand, v1.CapacityTypeSpot} {
				price := s.PriceOnDemand
				if ct == v1.CapacityTypeSpot {
					price *= SpotDiscount
				}
				// a zone surcharge far below any real price difference keeps every price distinct: OrderByPrice is an
				// unstable sort (types.go:223), ties would make the order depend on it
				price += float64(zi) * 1e-6
				offs = append(offs, &cloudprovider.Offering{
					Requirements: scheduling.NewRequirements(
						scheduling.NewRequirement(corev1.LabelTopologyZone, corev1.NodeSelectorOpIn, z),
						scheduling.NewRequirement(v1.CapacityTypeLabelKey, corev1.NodeSelectorOpIn, ct)),
					Price: price, Available: true,
				})
			}
		}
		out = append(out, &cloudprovider.InstanceType{
			Name: s.Name, Requirements: reqs, Offerings: offs,
			Capacity: corev1.ResourceList{
				corev1.ResourceCPU:              *resource.NewQuantity(s.VCPU, resource.DecimalSI),
				corev1.ResourceMemory:           *resource.NewQuantity(s.MemoryGiB<<30, resource.BinarySI),
				corev1.ResourceEphemeralStorage: *resource.NewQuantity(t.Overhead.EphemeralGiB<<30, resource.BinarySI),
				corev1.ResourcePods:             *resource.NewQuantity(t.Overhead.Pods, resource.DecimalSI),
				"nvidia.com/gpu":                *resource.NewQuantity(s.GPU, resource.DecimalSI),
			},
			Overhead: &cloudprovider.InstanceTypeOverhead{
				KubeReserved: corev1.ResourceList{
					corev1.ResourceCPU:    *resource.NewMilliQuantity(t.Overhead.KubeReservedCPUMilli, resource.DecimalSI),
					corev1.ResourceMemory: *resource.NewQuantity(t.Overhead.KubeReservedMemoryMiB<<20, resource.BinarySI),
				},
				EvictionThreshold: corev1.ResourceList{
					corev1.ResourceMemory: *resource.NewQuantity(t.Overhead.EvictionMemoryMiB<<20, resource.BinarySI),
				},
			},
		})
	}
	return out, nil
}
