"""b200sched — Python mirror of the host side of libb200sched (model building, synthetic inputs,
ctypes bindings of the C-ABI in include/b200sched.h).  The compute path is the CUDA library; this
package never falls back to a CPU implementation."""
