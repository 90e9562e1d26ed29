'''Developer script: wave size x stream count sweep of the Monte-Carlo trace (run under gpurun).
usage: gpu_wave_sweep.py [scene] [n_rays]   prints ms per request, best of 3'''
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from freecad.optics_design_workbench_b200 import engine
from freecad.optics_design_workbench_b200.simulation.setup import prepare
name = sys.argv[1] if len(sys.argv) > 1 else 'lensesAndMirrors'
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
eng = engine.Engine(0)
sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', name+'.npz'))
ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
cfg = sim.cfg(store_hits=True, hit_capacity=int(1.05*n)+1024)
with ds.trace_mc(dsrc, cfg, 0x0DDB1A5E, 0, n) as res:
  pass
print('streams \\ log2(wave): ' + ' '.join(f'{w:7d}' for w in range(16, 24)))
for streams in (1, 2, 3, 4, 6, 8):
  row = []
  for w in range(16, 24):
    os.environ['ODW_STREAMS'], os.environ['ODW_RAYS_PER_LAUNCH'] = str(streams), str(1 << w)
    best = 1e9
    for rep in range(3):
      with ds.trace_mc(dsrc, cfg, 0x0DDB1A5E, rep*n, n) as res:
        best = min(best, res.kernel_ms)
        segs = res.counts['segments']
    row.append(best)
  print(f'{streams:7d}               ' + ' '.join(f'{x:7.2f}' for x in row), flush=True)
print('segments per request', segs)
