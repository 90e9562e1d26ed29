'''Developer script: a few device-binned launches of a scene (for ncu).  usage: gpu_binned.py scene n_rays bins reps [nobins]'''
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from types import SimpleNamespace
from freecad.optics_design_workbench_b200 import engine
from freecad.optics_design_workbench_b200.simulation.setup import prepare
import bench
name, n, bins, reps = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3]), int(sys.argv[4])
nobins = len(sys.argv) > 5
eng = engine.Engine(0)
sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', name+'.npz'))
ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
cfg = sim.cfg(store_hits=False, binnings=[] if nobins else bench.binning_specs(sim, bins))
for rep in range(reps):
  with ds.trace_mc(dsrc, cfg, 0x0DDB1A5E, rep*n, n) as res:
    c, ms = res.counts, res.kernel_ms
  print(f'{name} n={n} bins={0 if nobins else bins}: {ms:.2f} ms {c["segments"]/ms*1e3:.3e} seg/s {n/ms*1e3:.3e} rays/s hits {c["hits"]}', flush=True)
