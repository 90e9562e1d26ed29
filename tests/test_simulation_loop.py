'''
Host logic around the engine on the CPU: runSimulation / runSimulationIteration / the hit writer, with the
oracle-backed test double standing in for the GPU engine (tests/oracle_engine.py).  The assertions follow the
reference's own integration tests where they exist (test/21-simulation-modes/run-simulations.py:42-64: a run
ends on EndAfterHits=1e3 with > 999 hits, on EndAfterRays=1e3 with > 100 hits).
'''
import glob
import os
import pickle
import re

import numpy as np
import pytest

from freecad.optics_design_workbench_b200.simulation import simulation_loop, results_store
from freecad.optics_design_workbench_b200.simulation.setup import prepare

from conftest import SCENES
from oracle_engine import OracleEngine


def load_hits(run_folder, pattern='**'):
  'what RawFolder._load does (reference jupyter_utils/freecad_document.py:1491-1504), restated with numpy only'
  result = {}
  for path in sorted(glob.glob(f'{run_folder}/{pattern}/*-hits.pkl', recursive=True)):
    with open(path, 'rb') as f:
      data = pickle.load(f)
    for k, v in data.items():
      if k not in result:
        result[k] = v
      elif isinstance(v, str):
        if isinstance(result[k], str) and result[k] != v:
          result[k] = [result[k], v]
        elif not isinstance(result[k], str) and v not in result[k]:
          result[k] = list(result[k])+[v]
      else:
        result[k] = np.concatenate([result[k], v], axis=0)
  return result


@pytest.fixture()
def engine():
  return OracleEngine()


def test_run_ends_after_rays_like_reference_test_21(tmp_path, engine):
  sim = prepare(os.path.join(SCENES, 'minimal.npz'))
  run = simulation_loop.runSimulation(sim, 'true', engine=engine, basePath=str(tmp_path/'minimal.OpticsDesign'),
                                      settings=dict(EndAfterRays=1e3), maxBatchRays=300)
  hits = load_hits(run)
  assert len(hits['points']) > 100                     # reference assertion
  # a single reference worker stops at the first ray count strictly above the limit: 11 iterations of 100 rays
  assert len(hits['points']) == 1100
  assert hits['points'].shape == (1100, 3) and hits['directions'].shape == (1100, 3)
  assert hits['powers'].dtype == np.float64 and hits['isEntering'].dtype.kind == 'i'
  assert hits['source'] == 'OpticalPointSource' and hits['obj'] == 'OpticalAbsorberGroup'
  np.testing.assert_allclose(hits['points'][:, 2], 15.0, atol=1e-9)      # entry face of the absorber box
  assert np.all(hits['isEntering'] == 1) and np.all(hits['powers'] == 1.0)


def test_run_ends_after_hits_like_reference_test_21(tmp_path, engine):
  sim = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
  run = simulation_loop.runSimulation(sim, 'true', engine=engine, basePath=str(tmp_path/'x.OpticsDesign'),
                                      settings=dict(EndAfterHits=1e3, EndAfterRays=np.inf), maxBatchRays=400)
  hits = load_hits(run)
  assert len(hits['points']) > 999


def test_result_tree_layout(tmp_path, engine):
  base = str(tmp_path/'lensesAndMirrors.OpticsDesign')
  sim = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
  run = simulation_loop.runSimulation(sim, 'singletrue', engine=engine, basePath=base)
  assert run == f'{base}/raw/simulation-run-000000'
  assert os.path.isdir(f'{base}/notebooks') and os.path.isfile(f'{base}/README.md')
  uid = [f for f in os.listdir(run) if f.startswith('uid-')]
  assert len(uid) == 1 and re.fullmatch(r'uid-[0-9a-f-]{36}', uid[0])       # RawFolder requires exactly one
  with open(f'{run}/global-info.pkl', 'rb') as f:
    info = pickle.load(f)
  assert set(info) == {'activeSimulationSettings', 'lightSources', 'opticalObjects'}
  assert info['lightSources'][0]['placementPathsAndMatrices'][0]['gpM'].shape == (4, 4)
  files = glob.glob(f'{run}/source-*/object-*/*-hits.pkl')
  assert files, os.listdir(run)
  for p in files:
    rel = os.path.relpath(p, run).split(os.sep)
    assert rel[0] == 'source-'+sim.source_records[0]['label'] and rel[1].startswith('object-')   # Labels, not Names
    assert re.fullmatch(r'\d+-pid\d+-thread\d+-hits\.pkl', rel[2])
    d = pickle.load(open(p, 'rb'))
    assert list(d)[:6] == ['source', 'obj', 'points', 'directions', 'powers', 'isEntering']
    assert d['source'] == sim.source_records[0]['name'] and d['obj'] in sim.scene.group_names        # Names inside the file
  assert not os.path.exists(f'{run}/progress')         # removed by cleanup like the reference
  # a second run gets the next index
  run2 = simulation_loop.runSimulation(sim, 'singletrue', engine=engine, basePath=base)
  assert run2.endswith('simulation-run-000001')


def test_fans_mode_with_store_hit_metadata(tmp_path, engine):
  sim = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
  sim.settings['store_hit_keys'] = ['InitPoint', 'InitDirection', 'InitPower', 'InitWavelength', 'InitPhi', 'InitTheta',
                                    'RayIndex', 'FanIndex', 'TotalFanCount', 'TotalRaysInFan']
  run = simulation_loop.runSimulation(sim, 'fans', engine=engine, basePath=str(tmp_path/'f.OpticsDesign'))
  hits = load_hits(run)
  n = len(hits['powers'])
  assert n == 40                                       # 2 fans x 20 rays, every ray ends on the absorber
  for k in results_store.HIT_METADATA_KEYS:
    assert k in hits and len(hits[k]) == n, k
  assert hits['initPoint'].shape == (n, 3) and hits['initDirection'].shape == (n, 3)
  assert set(hits['fanIndex']) == {0, 1} and np.all(hits['totalFanCount'] == 2) and np.all(hits['totalRaysInFan'] == 20)
  assert sorted(hits['rayIndex'][hits['fanIndex'] == 0]) == list(range(-9, 11)) or sorted(hits['rayIndex'][hits['fanIndex'] == 0]) == list(range(-10, 10))
  np.testing.assert_allclose(hits['initWavelength'], 500.0)


def test_monte_carlo_metadata_comes_from_the_same_philox_stream(tmp_path, engine):
  sim = prepare(os.path.join(SCENES, 'minimal.npz'))
  sim.settings['store_hit_keys'] = ['InitTheta', 'InitPhi', 'InitDirection']
  run = simulation_loop.runSimulation(sim, 'singletrue', engine=engine, basePath=str(tmp_path/'m.OpticsDesign'))
  hits = load_hits(run)
  # minimal scene known answer: hit = (15 tan(theta) sin(phi), -15 tan(theta) cos(phi), 15)
  th, ph = hits['initTheta'], hits['initPhi']
  want = np.stack([15*np.tan(th)*np.sin(ph), -15*np.tan(th)*np.cos(ph), np.full(len(th), 15.0)], axis=1)
  np.testing.assert_allclose(hits['points'], want, atol=1e-9)
  np.testing.assert_allclose(hits['directions'], hits['initDirection'], atol=1e-12)


def test_progress_files_and_master_summary(tmp_path, engine):
  sim = prepare(os.path.join(SCENES, 'minimal.npz'))
  run = simulation_loop.runSimulation(sim, 'true', engine=engine, basePath=str(tmp_path/'p.OpticsDesign'),
                                      settings=dict(EndAfterRays=500), maxBatchRays=200, keepProgressFiles=True)
  masters = sorted(glob.glob(f'{run}/progress/master-*'))
  assert masters
  last = pickle.load(open(masters[-1], 'rb'))
  assert set(last) == {'simulationType', 'totalIterations', 'totalTracedRays', 'totalRecordedHits', 'totalRecordedRays',
                       'endAfterIterations', 'endAfterRays', 'endAfterHits'}
  assert last['totalTracedRays'] == 600 and last['totalIterations'] == 6 and last['endAfterRays'] == 500


def test_writer_pads_missing_metadata_with_nan(tmp_path):
  st = results_store.SimulationResults('true', str(tmp_path/'w.OpticsDesign'))
  src, obj = ('S', 'S label'), ('G', 'G label')
  st.addRayHits(src, obj, np.zeros((2, 3)), np.ones((2, 3)), [1, 1], [1, 0], dict(initTheta=[0.1, 0.2]))
  st.addRayHits(src, obj, np.zeros((1, 3)), np.ones((1, 3)), [0.5], [1])
  st.flush()
  d = pickle.load(open(st.writtenFiles[0], 'rb'))
  assert 'source-S label/object-G label' in st.writtenFiles[0]
  assert d['source'] == 'S' and d['obj'] == 'G' and len(d['powers']) == 3
  assert np.isnan(d['initTheta'][2]) and d['initTheta'][0] == 0.1
  assert st.totalRecordedHits == 3


def test_lent_batches_are_written_in_parallel_slices_with_the_same_content(tmp_path, monkeypatch):
  '''
  addRayHits(..., borrowed=True) (large batches straight from the page-locked delivery buffers): the batch is cut into equally
  sized files written by a thread pool.  Whatever the slicing, the files hold the rows of the batch once, in order, with the
  reference's keys and dtypes (results_store.py:369-460), next to files of the buffered path.
  '''
  rng = np.random.default_rng(5)
  n = 10007
  pts, dirs, pw = rng.normal(size=(n, 3)), rng.normal(size=(n, 3)), rng.random(n)
  ent = (rng.random(n) > 0.5).astype(np.uint8)
  md = dict(initTheta=rng.random(n))
  src, obj = ('S', 'S label'), ('G', 'G label')
  for rows, threads in ((1 << 22, 32), (1000, 3), (999, 32), (n, 1)):
    monkeypatch.setattr(results_store.SimulationResults, 'ROWS_PER_FILE', rows)
    monkeypatch.setattr(results_store.SimulationResults, 'WRITER_THREADS', threads)
    monkeypatch.setattr(results_store.SimulationResults, 'DIRECT_WRITE_BYTES', 1000)     # small batches are copied into the buffer instead
    st = results_store.SimulationResults('true', str(tmp_path/f'w{rows}-{threads}.OpticsDesign'))
    st.addRayHits(src, obj, pts, dirs, pw, ent, md, borrowed=True)
    st.addRayHits(src, obj, pts[:3], dirs[:3], pw[:3], ent[:3], dict(initTheta=md['initTheta'][:3]))     # buffered path
    st.flush()
    assert st.totalRecordedHits == n + 3
    files = sorted(f for f in st.writtenFiles if f.endswith('-hits.pkl'))
    assert len(set(files)) == len(files)
    if rows < n:
      assert len(files) >= n//rows
      sizes = [len(pickle.load(open(f, 'rb'))['powers']) for f in files]
      assert max(sizes) <= rows
    got = load_hits(os.path.dirname(os.path.dirname(os.path.dirname(files[0]))))
    assert got['source'] == 'S' and got['obj'] == 'G'
    assert got['isEntering'].dtype == np.int64 and got['points'].dtype == np.float64
    order = np.lexsort((got['powers'], got['initTheta']))
    want = dict(points=np.concatenate([pts, pts[:3]]), directions=np.concatenate([dirs, dirs[:3]]), powers=np.concatenate([pw, pw[:3]]),
                isEntering=np.concatenate([ent, ent[:3]]).astype(np.int64), initTheta=np.concatenate([md['initTheta'], md['initTheta'][:3]]))
    worder = np.lexsort((want['powers'], want['initTheta']))
    for k, v in want.items():
      assert np.array_equal(got[k][order], v[worder]), (rows, threads, k)


def test_unknown_action_and_missing_end_criterion(tmp_path, engine):
  sim = prepare(os.path.join(SCENES, 'minimal.npz'))
  with pytest.raises(ValueError):
    simulation_loop.runSimulation(sim, 'stop', engine=engine, basePath=str(tmp_path/'a'))
  with pytest.raises(ValueError):
    simulation_loop.runSimulation(sim, 'true', engine=engine, basePath=str(tmp_path/'b'),
                                  settings=dict(EndAfterRays=np.inf, EndAfterHits=np.inf, EndAfterIterations=np.inf))


@pytest.mark.reference
def test_engine_bridge_with_reference_side_objects(tmp_path, engine):
  '''
  The stub of INTEGRATION.md: FreeCAD-side objects (here stand-ins with the attributes the bridge reads) drive the
  engine from the saved benchmark project, and the hits land in the reference store's run folder.
  '''
  import shutil
  import types
  from freecad.optics_design_workbench_b200 import engine_bridge
  fcstd = str(tmp_path/'minimal.FCStd')
  shutil.copy('/root/reference/benchmark/minimal.FCStd', fcstd)
  engine_bridge.set_engine_factory(lambda: engine)
  try:
    assert engine_bridge.available()
    obj = types.SimpleNamespace(Name='OpticalPointSource', Document=types.SimpleNamespace(FileName=fcstd))
    base = results_store.results_folder_path(fcstd)
    ref_store = types.SimpleNamespace(basePath=base, simulationRunFolder='raw/simulation-run-000000', simulationType='true',
                                      totalTracedRays=0, totalRecordedHits=0, flushEverySeconds=5)
    for _ in range(3):
      engine_bridge.run_iteration(None, obj, mode='true', store=ref_store)
    engine_bridge.flush(ref_store)
    assert ref_store.totalTracedRays == 300 and ref_store.totalRecordedHits == 300
    hits = load_hits(f'{base}/raw/simulation-run-000000')
    assert len(hits['points']) == 300 and hits['obj'] == 'OpticalAbsorberGroup'
  finally:
    engine_bridge.set_engine_factory(None)


def test_replay_source_re_emits_stored_hits(tmp_path, engine, oracle):
  '''
  ReplaySourceProxy (reference freecad_elements/replay_source.py:73-166): hits stored by one simulation are the rays of
  the next; the run ends when the stock is used up.
  '''
  # stage 1: rays of the point source stopped on a Vacuum plane in front of the optics
  sim = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
  sa = sim.source_args(0)
  s = oracle.sample_mc(sa, 7, 0, 1000)
  stage1 = results_store.SimulationResults('true', str(tmp_path/'stage1.OpticsDesign'))
  pts = s['origins'] + s['directions']*5.0                                  # 5 mm down the beam
  stage1.addRayHits(('src', 'src'), ('Plane', 'Plane'), pts, s['directions'], np.full(1000, 0.7), np.ones(1000))
  stage1.flush()
  # stage 2: a replay source pointing at stage 1's run folder, shifted by its own placement
  gpM = np.eye(4); gpM[:3, 3] = [0, 0, 0.25]
  sim2 = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
  sim2.source_records = [dict(name='OpticalReplaySource', label='replay', proxy='ReplaySourceProxy', source_id=0, gpM=gpM,
                              ignored=[], ReplayFromDir=stage1.runFolderPath(), RaysPerIterationScale=1.0,
                              MaxRayLengthScale=1.0, MaxIntersectionsScale=1.0, Wavelength=500.0)]
  run = simulation_loop.runSimulation(sim2, 'true', engine=engine, basePath=str(tmp_path/'stage2.OpticsDesign'),
                                      settings=dict(RaysPerIteration=300), maxBatchRays=300)
  hits = load_hits(run)
  want = oracle.trace_rays(sim2.scene, sim2.cfg(wavelength=1.0), pts+[0, 0, 0.25], s['directions'], np.full(1000, 0.7))
  assert len(hits['points']) == want['counts']['hits'] > 900
  key = lambda p: np.lexsort(np.round(p, 9).T)
  np.testing.assert_allclose(hits['points'][key(hits['points'])], want['hits']['points'][key(want['hits']['points'])], atol=1e-9)
  np.testing.assert_allclose(hits['powers'], 0.7*np.ones(len(hits['powers'])))       # mirrors have Reflectivity 1 here
  assert hits['source'] == 'OpticalReplaySource'


def test_record_rays_writes_the_reference_rays_file(tmp_path, engine):
  '''
  RecordRays (reference generic_source.py:80-100, results_store.py:241-257,380-403): *-rays.pkl = list of dict(points (M+1,3),
  powers (M,), media [Name | None]*M) per ray, next to the hit files of the RecordHits groups.
  '''
  sim = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
  sim.source_records[0]['RecordRays'] = True
  run = simulation_loop.runSimulation(sim, 'singletrue', engine=engine, basePath=str(tmp_path/'r.OpticsDesign'))
  files = glob.glob(f'{run}/source-*/*-rays.pkl')
  assert len(files) == 1 and os.path.dirname(files[0]).endswith('source-'+sim.source_records[0]['label'])
  rays = pickle.load(open(files[0], 'rb'))
  assert len(rays) == 100
  full = [r for r in rays if len(r['powers']) == 7]
  assert len(full) > 90                                  # the 7-segment path of SURVEY Appendix B
  r = full[0]
  assert r['points'].shape == (8, 3) and len(r['media']) == 7
  np.testing.assert_allclose(r['points'][0], 0.0)        # the source sits at the origin
  lens = [n for n, g in zip(sim.scene.group_names, sim.scene.groups) if int(g['optical_type']) == 1][0]
  assert r['media'] == [None, None, lens, None, lens, None, None]
  np.testing.assert_allclose(r['powers'], 1.0)
  assert abs(r['points'][-1][2]-73.0) < 1e-6             # ends on the absorber box (z = 73)
  # hit files still hold only the RecordHits groups
  hits = load_hits(run)
  assert len(hits['points']) >= 90 and set(np.atleast_1d(hits['obj']).tolist()) <= set(sim.scene.group_names)
  # fans mode records rays too
  run2 = simulation_loop.runSimulation(sim, 'fans', engine=engine, basePath=str(tmp_path/'r.OpticsDesign'))
  rays2 = pickle.load(open(glob.glob(f'{run2}/source-*/*-rays.pkl')[0], 'rb'))
  assert len(rays2) == 40


def test_flag_files_cancel_a_running_simulation(tmp_path, engine):
  'the reference protocol: simulation-is-running while it runs, simulation-is-canceled / -done stop it between batches'
  import threading
  base = str(tmp_path/'c.OpticsDesign')
  sim = prepare(os.path.join(SCENES, 'minimal.npz'))
  seen = {}
  def canceller():
    import time
    for _ in range(2000):
      if simulation_loop.query_status(base, 'simulation-is-running'):
        seen['running'] = True
        simulation_loop.cancelSimulation(base)
        return
      time.sleep(0.005)
  t = threading.Thread(target=canceller); t.start()
  run = simulation_loop.runSimulation(sim, 'true', engine=engine, basePath=base, settings=dict(EndAfterRays=1e9), maxBatchRays=20000)
  t.join()
  assert seen.get('running') and not simulation_loop.query_status(base, 'simulation-is-running')
  assert simulation_loop.query_status(base, 'simulation-is-canceled')
  n = len(load_hits(run)['points'])
  assert 0 < n < 1e8                                      # stopped long before EndAfterRays; what was traced is on disk
  # a run that reaches its end criterion marks itself done
  simulation_loop.runSimulation(sim, 'true', engine=engine, basePath=base, settings=dict(EndAfterRays=500), maxBatchRays=200)
  assert simulation_loop.query_status(base, 'simulation-is-done') and not simulation_loop.query_status(base, 'simulation-is-canceled')


def test_pseudo_mode_thins_towards_its_expected_histogram(tmp_path, engine):
  '''
  'pseudo' / 'singlepseudo' (reference point_source.py:671-679 -> drawPseudo, random_number_generator.py:562-682).  The
  procedure deletes samples from the histogram bin that exceeds its expected share (density at the bin centre, a coarse
  int((f sqrt(iterations) N)**(1/6))-bin grid) the most; so measured with ITS criterion the result is closer to the
  expectation than true random draws of the same size.  (With the default 100 rays per iteration that grid is 2x2 — the
  mode exists for nicer looking displayed rays, not for accuracy.)
  '''
  import sympy as sy
  from freecad.optics_design_workbench_b200.distributions import sampler_tables as st
  sim = prepare(os.path.join(SCENES, 'hugeArray.npz'))             # exp(-theta^2/0.2^2): a wide beam
  rec = sim.source_records[0]
  tables = sim.source_args(0).tables
  expr, _ = st.point_source_density(rec['PowerDensity'], float(rec['FocalLength']))
  lam = sy.lambdify([sy.Symbol('phi'), sy.Symbol('theta')], expr, modules=['numpy'])
  rng = np.random.default_rng(5)
  N, bins = 4000, 3                                                # int((0.1*sqrt(50)*4000)**(1/6)) = 3
  def misfit(theta, phi):
    hist, edges = np.histogramdd(np.stack([theta, phi]).T, bins=bins)
    c = [(e[1:]+e[:-1])/2 for e in edges]
    expected = lam(*np.meshgrid(*reversed(c)))
    return np.abs(hist/hist.sum()-expected/expected.sum()).max()
  th, ph = st.draw_pseudo(tables, expr, N, rng)
  assert len(th) == len(ph) == N and th.min() >= 0 and th.max() <= np.pi/4 and ph.min() >= 0 and ph.max() <= 2*np.pi
  t2, p2 = tables.draw_from_uniforms(rng.random(N), rng.random(N))
  assert misfit(th, ph) < misfit(t2, p2)
  a = st.draw_pseudo(tables, expr, 100, np.random.default_rng(1))
  b = st.draw_pseudo(tables, expr, 100, np.random.default_rng(1))
  np.testing.assert_array_equal(a[0], b[0])                        # reproducible for a given stream
  run = simulation_loop.runSimulation(prepare(os.path.join(SCENES, 'minimal.npz')), 'singlepseudo', engine=engine,
                                      basePath=str(tmp_path/'p.OpticsDesign'))
  assert len(load_hits(run)['points']) == 100
  run = simulation_loop.runSimulation(prepare(os.path.join(SCENES, 'minimal.npz')), 'pseudo', engine=engine,
                                      basePath=str(tmp_path/'p.OpticsDesign'), settings=dict(EndAfterRays=250), maxBatchRays=200)
  assert len(load_hits(run)['points']) == 300


def test_draw_hands_the_traced_polylines_to_a_drawing_back_end(tmp_path, engine):
  '''
  draw=True (reference generic_source.py:102-138): one Part line per segment, in the light source's local frame, grouped into
  RaySegment features of the source.  Here the traced polylines go to a drawing back end; the recording one stands in for
  FreeCAD.  Fans mode (the GUI's default action) and a single Monte-Carlo iteration.
  '''
  from freecad.optics_design_workbench_b200.freecad_elements import ray_drawing
  from freecad.optics_design_workbench_b200.freecad_elements.generic_source import GenericSourceProxy
  sim = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
  ctx = simulation_loop.SimulationContext(sim, engine, seed=3)
  src = GenericSourceProxy(ctx, 0)
  backend = ray_drawing.RecordingBackend()
  counts = src.runSimulationIteration(mode='fans', draw=True, drawBackend=backend)
  assert backend.cleared == 1 and len(backend.rays) == 40 == counts['rays']
  assert sum(len(r) for r in backend.rays) == counts['segments']
  full = [r for r in backend.rays if len(r) == 7][0]
  np.testing.assert_allclose(full[0][0], 0.0, atol=1e-12)                  # starts at the source
  np.testing.assert_allclose(full[1:, 0], full[:-1, 1])                   # segments chain: end of one = start of the next
  assert abs(full[-1][1][2]-73.0) < 1e-6                                  # ends on the absorber (the source frame is the world frame here)
  counts = src.runSimulationIteration(mode='true', draw=True, drawBackend=backend)
  assert backend.cleared == 2 and len(backend.rays) == 100 == counts['rays']
  # a source with a placement: lines are drawn in its LOCAL frame (gpMi * p)
  sim.source_records[0]['gpM'] = np.array([[1, 0, 0, 5.0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
  ctx2 = simulation_loop.SimulationContext(sim, engine, seed=3)
  GenericSourceProxy(ctx2, 0).runSimulationIteration(mode='fans', draw=True, drawBackend=backend)
  np.testing.assert_allclose(backend.rays[0][0][0], 0.0, atol=1e-12)      # the ray starts at the source's own origin
  with pytest.raises(RuntimeError, match='FreeCAD is not importable'):
    src.runSimulationIteration(mode='fans', draw=True)
