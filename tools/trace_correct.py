import sys, time, os
sys.path.insert(0,'.')
import numpy as np, rattle_b200
from tools import synth
print('cpus', os.cpu_count())
rs = synth.config2(n_genes=int(sys.argv[1])).sorted_by_length()[0]
ctx = rattle_b200.Context(0)
cl = ctx.cluster_reads(rs.bases, rs.offsets, is_rna=False)
for it in range(2):
    t=time.time(); out = ctx.correct_reads(rs.bases, rs.quals, rs.offsets, cl); dt=time.time()-t
    print('correct %.3f s'%dt, {k:v for k,v in ctx.stats().items() if k.startswith('poa') or k=='total_ms'})
