// select.go — the vmSize pick of Provider.Create (pkg/providers/instance/instance.go:76-95), replacing
// `instanceTypes[0]` of Requirement.Values() — an unsorted list out of a Go map, i.e. nondeterministic as soon as the
// NodeClaim names more than one instance type (SURVEY.md fact 1) — by the cheapest entry under the catalog's prices, the
// selection the scheduler computed (OrderByPrice(reqs)[0], types.go:221-240).  ToNodeClaim writes the requirement as a
// name-sorted set (nodeclaimtemplate.go:79-82, requirement.go:140), so the price order has to be re-established here.
//
// Drop-in: in instance.go replace
//
//	vmSize := instanceTypes[0]
//
// by
//
//	vmSize, err := p.selectVMSize(nodeClaim, instanceTypes)
//	if err != nil { return err }
//
// NOT COMPILED IN THE AUTHORING IMAGE; b200s_create_select (csrc/encoder.cpp) is the same rule behind the C-ABI and is
// pinned by the reference's own test vectors (tests/test_oracle.py::test_create_select_reference_vectors).
package instance

import (
	"fmt"
	"math"

	corev1 "k8s.io/api/core/v1"
	karpenterv1 "sigs.k8s.io/karpenter/pkg/apis/v1"
	"sigs.k8s.io/karpenter/pkg/cloudprovider"
	"sigs.k8s.io/karpenter/pkg/scheduling"
)

// PriceSource is the catalog (go/pkg/providers/instancetype.Provider satisfies it).
type PriceSource interface {
	List() ([]*cloudprovider.InstanceType, error)
}

// selectVMSize returns the cheapest of the NodeClaim's instance types under its own requirements (zone, capacity type):
// min over compatible available offerings, exactly OrderByPrice's key; a name the catalog does not know keeps the
// reference's behaviour for a single-valued list and is an error otherwise.
func (p *Provider) selectVMSize(nodeClaim *karpenterv1.NodeClaim, instanceTypes []string) (string, error) {
	if len(instanceTypes) == 1 || p.prices == nil {
		return instanceTypes[0], nil
	}
	its, err := p.prices.List()
	if err != nil {
		return "", err
	}
	byName := map[string]*cloudprovider.InstanceType{}
	for _, it := range its {
		byName[it.Name] = it
	}
	reqs := scheduling.NewNodeSelectorRequirementsWithMinValues(nodeClaim.Spec.Requirements...)
	best, bestPrice := "", math.MaxFloat64
	for _, name := range instanceTypes {
		it, ok := byName[name]
		if !ok {
			return "", fmt.Errorf("instance type %q is not in the catalog", name)
		}
		price := math.MaxFloat64
		if ofs := it.Offerings.Available().Compatible(reqs); len(ofs) > 0 {
			price = ofs.Cheapest().Price
		}
		if price < bestPrice || best == "" {
			best, bestPrice = name, price
		}
	}
	_ = corev1.LabelInstanceTypeStable
	return best, nil
}
