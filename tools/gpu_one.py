'''Developer script: a few launches of one configuration (for ncu).  usage: gpu_one.py scene n_rays store(0/1) reps'''
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from freecad.optics_design_workbench_b200 import engine
from freecad.optics_design_workbench_b200.simulation.setup import prepare
name, n, store, reps = sys.argv[1], int(float(sys.argv[2])), bool(int(sys.argv[3])), int(sys.argv[4])
eng = engine.Engine(0)
sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', name+'.npz'))
ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
cfg = sim.cfg(store_hits=store, hit_capacity=int(1.05*n)+1024)
for rep in range(reps):
  with ds.trace_mc(dsrc, cfg, 0x0DDB1A5E, rep*n, n) as res:
    c, ms = res.counts, res.kernel_ms
  print(f'{name} n={n} store={store}: {ms:.2f} ms {c["segments"]/ms*1e3:.3e} seg/s', flush=True)
