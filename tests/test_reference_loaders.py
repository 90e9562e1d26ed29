'''
The result tree written by our hit writer, read back through the REFERENCE's own loaders
(jupyter_utils.RawFolder.loadHits / loadGlobalInfo, simulation.findPathsAndSanitize, io.unpickle, Hits,
Histogram) imported from /root/reference.  Runs in the build container only (marker `reference`).
'''
import os

import numpy as np
import pytest

from freecad.optics_design_workbench_b200.simulation import simulation_loop
from freecad.optics_design_workbench_b200.simulation.setup import prepare

from conftest import SCENES
from oracle_engine import OracleEngine

pytestmark = pytest.mark.reference


@pytest.fixture(scope='module')
def ref():
  import reference_shim
  return reference_shim.load()


def test_rawfolder_loads_our_tree(tmp_path, ref):
  sim = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
  sim.settings['store_hit_keys'] = ['InitTheta', 'InitPhi']
  run = simulation_loop.runSimulation(sim, 'true', engine=OracleEngine(), basePath=str(tmp_path/'l.OpticsDesign'),
                                      settings=dict(EndAfterRays=2000), maxBatchRays=700, flushEverySeconds=0)
  assert ref.RawFolder is not None
  folder = ref.RawFolder(run, timeout=1)
  tree = folder.tree()
  assert any(k.startswith('source-') for k in tree)
  info = folder.loadGlobalInfo()
  assert len(info['lightSources']) == 1 and len(info['opticalObjects']) == 4
  hits = folder.loadHits('*')                       # several files (one per flush): concatenated by the reference
  assert isinstance(hits, ref.hits.Hits)
  d = hits._hits if hasattr(hits, '_hits') else hits.__dict__
  n = 2100
  pts = hits.points() if callable(getattr(hits, 'points', None)) else None
  raw = folder._load('*', 'hits')
  assert raw['points'].shape[1] == 3 and len(raw['points']) == len(raw['powers']) == len(raw['initTheta'])
  assert abs(len(raw['points'])-n) <= 5             # ~1 recorded hit per ray (a few miss the absorber)
  assert raw['source'] == sim.source_records[0]['name']
  # sub-pattern selection like RawFolder.loadHits('source-*/object-*')
  sub = folder._load('source-*/object-*', 'hits')
  assert len(sub['points']) == len(raw['points'])


def test_device_binning_equals_reference_histogram_of_the_hit_list(tmp_path, ref):
  '''
  odw_result_histogram semantics are pinned against the reference's post-hoc path: Hits.histogram ->
  planeProject3dPoints -> Histogram -> numpy.histogram2d (jupyter_utils/hits.py:62-94,176-193, histogram.py:24-56)
  with an explicit plane, origin and bin edges.  The binned counts come from the oracle here (the CUDA kernel is
  checked against the oracle's bins in tests/test_gpu_parity.py).
  '''
  from oracle import Oracle
  sim = prepare(os.path.join(SCENES, 'minimal.npz'))
  run = simulation_loop.runSimulation(sim, 'true', engine=OracleEngine(), basePath=str(tmp_path/'m.OpticsDesign'),
                                      settings=dict(EndAfterRays=19999, RaysPerIteration=20000), maxBatchRays=1 << 20)
  hits = ref.RawFolder(run, timeout=1).loadHits('*')
  nu, nv, R = 24, 16, 0.5
  edges_u, edges_v = np.linspace(-R, R, nu+1), np.linspace(-0.8*R, 0.8*R, nv+1)
  hist = hits.histogram(planeNormal=np.array([0., 0., 1.]), xInPlaneVec=np.array([1., 0., 0.]),
                        origin=np.array([0., 0.]), bins=[edges_u, edges_v])
  assert hist.hist.shape == (nu, nv) and hist.hist.sum() > 15000
  # the same 20000 rays (Philox stream of seed/source 0, rays 0..19999) binned by the engine rule
  binning = dict(group=0, nu=nu, nv=nv, origin=(0, 0, 0), uaxis=(1, 0, 0), vaxis=(0, 1, 0),
                 u_range=(-R, R), v_range=(-0.8*R, 0.8*R))
  cfg = sim.cfg(store_hits=False, binnings=[binning])
  r = Oracle().trace_mc(sim.scene, sim.source_args(0), cfg, simulation_loop.DEFAULT_SEED, 0, 20000)
  assert np.array_equal(r['histograms'][0], hist.hist)
