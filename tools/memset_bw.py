import sys, os
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,'gpu-provisioner_b200'))
import __graft_entry__ as g; g.build()
from b200sched import capi
ctx=capi.Context(0); e0,e1=ctx.event(),ctx.event()
for n in (256<<20, 1<<30):
    for i in range(3):
        ctx.record(e0); ctx.flush_l2(n); ctx.record(e1)
        ms=ctx.elapsed_ms(e0,e1)
    print(f"memset {n>>20} MiB: {ms:.4f} ms -> {n/ms/1e6:.1f} GB/s write-only")
ctx.close()
