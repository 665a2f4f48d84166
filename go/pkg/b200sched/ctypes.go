package b200sched

/*
#include "b200sched.h"
*/
import "C"

import (
	"unsafe"

	corev1 "k8s.io/api/core/v1"
)

// small conversions shared by flatten.go / results.go (cgo types cannot be named outside files that import "C")

func C_int32(v int32) C.int32_t { return C.int32_t(v) }
func C_int64(v int64) C.int64_t { return C.int64_t(v) }

func opCode(op corev1.NodeSelectorOperator) C.int32_t {
	switch op {
	case corev1.NodeSelectorOpIn:
		return C.B200S_OP_IN
	case corev1.NodeSelectorOpNotIn:
		return C.B200S_OP_NOTIN
	case corev1.NodeSelectorOpExists:
		return C.B200S_OP_EXISTS
	case corev1.NodeSelectorOpDoesNotExist:
		return C.B200S_OP_DOESNOTEXIST
	case corev1.NodeSelectorOpGt:
		return C.B200S_OP_GT
	default:
		return C.B200S_OP_LT
	}
}

func effectCode(e corev1.TaintEffect) C.int32_t {
	switch e {
	case corev1.TaintEffectNoSchedule:
		return C.B200S_EFFECT_NOSCHEDULE
	case corev1.TaintEffectPreferNoSchedule:
		return C.B200S_EFFECT_PREFERNOSCHEDULE
	case corev1.TaintEffectNoExecute:
		return C.B200S_EFFECT_NOEXECUTE
	default:
		return C.B200S_EFFECT_NONE
	}
}

// unsafeIndex returns &base[i] for a C int32 array.
func unsafeIndex(base *C.int32_t, i int) unsafe.Pointer {
	return unsafe.Add(unsafe.Pointer(base), uintptr(i)*4)
}
