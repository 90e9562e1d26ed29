'''
Small traces for compute-sanitizer (memcheck / racecheck / synccheck), SURVEY.md §5: the warp-aggregated hit append and
device binning of the register-resident kernel (lensesAndMirrors, 2^16 rays), the chunked host delivery, and the
wavefront kernels with their ballot compaction (hugeArray, 2^18 rays).
  compute-sanitizer --tool memcheck  python tools/sanitize_case.py
  compute-sanitizer --tool racecheck python tools/sanitize_case.py
'''
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from freecad.optics_design_workbench_b200 import engine, _abi
from freecad.optics_design_workbench_b200.simulation.setup import prepare

SEED = 0x0DDB1A5E
eng = engine.Engine(0)
sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'lensesAndMirrors.npz'))
ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
n = 1 << 16
absorber = sim.scene.group_names.index('Absorber') if 'Absorber' in sim.scene.group_names else len(sim.scene.group_names)-1
binning = dict(group=absorber, nu=64, nv=64, origin=(-68.858, 0.0, 73.0), uaxis=(1, 0, 0), vaxis=(0, 1, 0), u_range=(-2, 2), v_range=(-2, 2))
with ds.trace_mc(dsrc, sim.cfg(store_hits=True, hit_capacity=2*n, binnings=[binning]), SEED, 0, n) as res:
  c = res.counts
  print('lensesAndMirrors', c, 'binned', res.histogram(0).sum())
os.environ['ODW_HOST_CHUNK'] = '20000'; os.environ['ODW_RAYS_PER_LAUNCH'] = '4096'
arrays = _abi.HitArrays(2*n)
counts, got = ds.trace_mc_host(dsrc, sim.cfg(), SEED, 0, n, arrays.view)
print('host delivery', counts, got)
del os.environ['ODW_HOST_CHUNK'], os.environ['ODW_RAYS_PER_LAUNCH']
ds.close(); dsrc.close()
sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'hugeArray.npz'))
ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
n = 1 << 18
with ds.trace_mc(dsrc, sim.cfg(store_hits=True, hit_capacity=2*n), SEED, 0, n) as res:
  print('hugeArray', res.counts)
ds.close(); dsrc.close()
eng.close()
print('done')
