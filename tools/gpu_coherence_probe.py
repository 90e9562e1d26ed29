'''Developer script: does ray order matter on the BVH path?  The Monte-Carlo rays of hugeArray traced as an explicit list in
draw order, sorted by direction (Morton code of the octahedral map), and shuffled.  (run under gpurun)'''
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from freecad.optics_design_workbench_b200 import engine
from freecad.optics_design_workbench_b200.simulation.setup import prepare

def morton2(a, b):
  def spread(x):
    x = x.astype(np.uint64)
    x = (x | (x << 16)) & 0x0000FFFF0000FFFF
    x = (x | (x << 8)) & 0x00FF00FF00FF00FF
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0F
    x = (x | (x << 2)) & 0x3333333333333333
    x = (x | (x << 1)) & 0x5555555555555555
    return x
  return spread(a) | (spread(b) << 1)

name = sys.argv[1] if len(sys.argv) > 1 else 'hugeArray'
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4_000_000
eng = engine.Engine(0)
sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', name+'.npz'))
ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
s = dsrc.sample(0x0DDB1A5E, 0, n)
o, d = s['origins'], s['directions']
dn = d/np.linalg.norm(d, axis=1)[:, None]
l1 = np.abs(dn).sum(axis=1)
px, py = dn[:, 0]/l1, dn[:, 1]/l1                     # octahedral map (upper/lower hemisphere folded by the sign of z)
neg = dn[:, 2] < 0
qx = np.where(neg, (1-np.abs(py))*np.sign(px), px); qy = np.where(neg, (1-np.abs(px))*np.sign(py), py)
key = morton2(((qx+1)*0.5*65535).astype(np.uint32), ((qy+1)*0.5*65535).astype(np.uint32))
orders = {'draw order': np.arange(n), 'sorted by direction': np.argsort(key, kind='stable'), 'shuffled': np.random.default_rng(1).permutation(n)}
cfg = sim.cfg(store_hits=True, hit_capacity=int(1.05*n)+1024)
for label, idx in orders.items():
  oo, dd = np.ascontiguousarray(o[idx]), np.ascontiguousarray(d[idx])
  best = 1e9
  for rep in range(3):
    with ds.trace_rays(cfg, oo, dd) as res:
      best = min(best, res.kernel_ms); c = res.counts
  print(f'{name} {n} rays, {label}: {best:.2f} ms, {c["segments"]/best*1e3:.3e} segments/s ({c["segments"]} segments)', flush=True)
