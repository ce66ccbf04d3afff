"""debug helper (torchrun, 2 ranks): dump what the fixed-capacity routing produced"""
import os, sys
import numpy, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
env = bench.Env()
from vmad_b200.pm import ParticleMesh
P, rank = env.world, env.rank
N, L = 32, 100.0
pm = ParticleMesh([N] * 3, L, comm=env.comm)
rng = numpy.random.default_rng(11)
x_all = rng.uniform(-0.3 * L, 1.3 * L, size=(20000, 3))
x = x_all[rank::P]
lay0 = pm.decompose(x)
plan = pm.plan_static_routing()
lay = pm.decompose(x)
torch.cuda.synchronize()
idx = lay.indices.cpu().numpy()
rc = lay.rcnt.cpu().numpy()
nvalid = int(rc.sum())
rb = numpy.array(plan.rbase)
out = open(os.path.join(ROOT, 'gpurun_out', 'debug_static_rank%d.txt' % rank), 'w')
print("cap", plan.cap.tolist(), "rbase", plan.rbase, "sbase", plan.sbase, "rcnt", rc.tolist(), "scnt", lay.scnt.cpu().tolist(), file=out)
print("sentinel", pm._sentinel(), "mesh start0 %d nloc0 %d ghost %d" % (pm._mesh.start0, pm._mesh.nloc0, pm._mesh.ghost), file=out)
pre = idx[:nvalid]
neg = (pre < 0).sum()
src = numpy.searchsorted(rb, numpy.maximum(pre, 0), side='right') - 1
k = pre - rb[src]
gap = (k >= rc[src]) | (pre < 0)
print("prefix: %d slots, %d negative, %d gap rows" % (nvalid, neg, gap.sum()), file=out)
xs = lay._xs.cpu().numpy()
keys = pm.cell_keys(lay._xs).cpu().numpy() if False else None
print("first gap slots", numpy.nonzero(gap)[0][:10].tolist(), file=out)
for s in numpy.nonzero(gap)[0][:10]:
    print(" slot", s, "recv row", pre[s], "src", src[s], "k", k[s], "xs", xs[s].tolist(), file=out)
tail = idx[nvalid:]
print("tail: %d slots, %d marked" % (len(tail), (tail < 0).sum()), file=out)
print("xs of last 3 slots", xs[-3:].tolist(), file=out)
# sentinel rows anywhere in xs?
sent = numpy.array(pm._sentinel())
is_sent = (numpy.abs(xs - sent).max(axis=1) == 0)
print("sentinel rows in xs: %d (expected %d), first at slot %s" % (is_sent.sum(), len(idx) - nvalid, numpy.nonzero(is_sent)[0][:5].tolist()), file=out)
out.close()
env.dist.barrier()
env.dist.destroy_process_group()
