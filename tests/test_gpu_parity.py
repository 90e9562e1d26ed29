'''
Parity tests proper (-m gpu): the CUDA path, called through the C ABI (libodw_b200.so via ctypes),
against the CPU oracle on the same seeded inputs; plus size-independent properties at full benchmark size.

Bars: object/face sequence per ray identical (integer work: exact).  Hit positions and directions: the north
star allows max(tessellation deflection, 1e-6 relative); everything here is closed form, so the tests demand
1e-8 mm absolute on positions (1e-10 relative at the 100 mm scene scale) and 1e-9 on direction components.
Scenes made of many small spheres (hugeArray, the BVH test) are chaotic billiards: a 1e-16 rounding difference
grows ~10x per bounce, so positions are compared up to a stated bounce depth there and deeper bounces are held
to the sequence check only.  Rays within tolerance of a face edge are exempt by the north star: at most 0.05 % of
the rays of such a scene may differ in sequence, everywhere else none.
'''
import ctypes as C
import os

import numpy as np
import pytest

from freecad.optics_design_workbench_b200 import _abi, engine
from freecad.optics_design_workbench_b200.scene_export import primitives as prim
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder

pytestmark = pytest.mark.gpu
SEED = 0x0DDB1A5E
POS_TOL, DIR_TOL = 1e-8, 1e-9
SCENES = ['minimal', 'lensesAndMirrors', 'lensesAndMirrorsSequential', 'hugeArray']
# sphere billiards: near-grazing hits amplify rounding by 1/sqrt(discriminant) per bounce; still 10x inside the
# north star's 1e-6 relative bar at the 100 mm scene scale
CHAOTIC = dict(max_bounce=2, pos_tol=1e-5, dir_tol=1e-5)


def per_ray_sequences(h, n):
  seq = [[] for _ in range(n)]
  for r, f in zip(h['ray_index'], h['face_id']):
    seq[int(r)].append(int(f))
  return seq


def compare_hits(g, o, n, max_bad_fraction=0.0, base=0, max_bounce=None, pos_tol=POS_TOL, dir_tol=DIR_TOL):
  'g, o: sorted hit dicts of GPU and oracle'
  if (len(g['face_id']) == len(o['face_id']) and np.array_equal(g['face_id'], o['face_id'])
      and np.array_equal(g['ray_index'], o['ray_index'])):
    keep = np.ones(len(g['face_id']), dtype=bool)
    bad = 0
  else:
    sg = per_ray_sequences({**g, 'ray_index': g['ray_index']-base}, n)
    so = per_ray_sequences({**o, 'ray_index': o['ray_index']-base}, n)
    bad_rays = {i for i in range(n) if sg[i] != so[i]}
    bad = len(bad_rays)
    assert bad <= max_bad_fraction*n, f'{bad} of {n} rays have different face sequences'
    mg = ~np.isin(g['ray_index']-base, list(bad_rays))
    mo = ~np.isin(o['ray_index']-base, list(bad_rays))
    g = {k: v[mg] for k, v in g.items()}
    o = {k: v[mo] for k, v in o.items()}
  assert np.array_equal(g['group'], o['group']) and np.array_equal(g['bounce'], o['bounce'])
  assert np.array_equal(g['is_entering'], o['is_entering'])
  assert np.array_equal(g['medium'], o['medium'])          # medium the segment ending at the hit ran through (rays.pkl `media`)
  sel = slice(None) if max_bounce is None else (g['bounce'] <= max_bounce)
  assert np.abs(g['points'][sel]-o['points'][sel]).max(initial=0) < pos_tol
  assert np.abs(g['directions'][sel]-o['directions'][sel]).max(initial=0) < dir_tol
  assert np.abs(g['powers']-o['powers']).max(initial=0) < 1e-14
  return bad


@pytest.mark.parametrize('name', SCENES)
def test_sampler_matches_oracle(name, gpu_engine, oracle, sims):
  sim = sims(name)
  sa = sim.source_args(0)
  dsrc = gpu_engine.source(sa)
  n = 100000
  g = dsrc.sample(SEED, 5_000_000_000, n)          # counters beyond 2^32 exercise the 64-bit Philox counter
  o = oracle.sample_mc(sa, SEED, 5_000_000_000, n)
  assert np.abs(g['first']-o['first']).max() < 1e-15
  assert np.abs(g['phi']-o['phi']).max() < 1e-14
  assert np.abs(g['origins']-o['origins']).max() < 1e-12
  assert np.abs(g['directions']-o['directions']).max() < 1e-14


def test_sampler_phi_dependent_tables(gpu_engine, oracle):
  'conditional rows (n_rows = n_phi-1) + nearest-row rule, incl. the collimated (r, phi) kind'
  from freecad.optics_design_workbench_b200.distributions import build_tables
  for expr, var, dom, kind in (('(exp(-theta**2/0.05)*(1+0.8*cos(phi)**2))*abs(sin(theta))', 'theta', (0, np.pi/3), 0),
                               ('(exp(-r**2/4)*(2+sin(phi)))*abs(r)', 'r', (0, 6), 1)):
    t = build_tables(expr, var, dom, (0.3, 5.1), 2001, 41)
    m = prim.translation(1, 2, 3) @ prim.rotation((1, 1, 0), 0.4)
    sa = _abi.SourceArgs(t, kind=kind, source_id=3, gpM=m, focal_length=0.0 if kind else 2.5)
    g = gpu_engine.source(sa).sample(11, 0, 50000)
    o = oracle.sample_mc(sa, 11, 0, 50000)
    for k in ('first', 'phi', 'origins', 'directions'):
      assert np.abs(g[k]-o[k]).max() < 1e-13, k


@pytest.mark.parametrize('name', SCENES)
def test_monte_carlo_hits_match_oracle(name, gpu_engine, oracle, sims):
  'every intersection recorded: per-ray face sequence, positions, directions, powers, isEntering'
  sim = sims(name)
  n = 30000
  cap = n*(int(sim.settings['MaxIntersections']) if name == 'hugeArray' else 10)
  cfg = sim.cfg(record_all_hits=True, hit_capacity=cap)
  sa = sim.source_args(0)
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sa)
  with ds.trace_mc(dsrc, cfg, SEED, 1000, n) as res:
    gc, gh = res.counts, res.hits(sort=True)
  o = oracle.trace_mc(sim.scene, sa, cfg, SEED, 1000, n, hit_capacity=cap, threads=0)
  bad = compare_hits(gh, o['hits'], n, max_bad_fraction=5e-4 if name == 'hugeArray' else 0.0, base=1000,
                     **(CHAOTIC if name == 'hugeArray' else {}))
  if bad == 0:
    assert gc == o['counts']
  else:
    assert abs(gc['segments']-o['counts']['segments']) <= 100*bad


@pytest.mark.parametrize('name', ['grating', 'gaussian', 'gettingStarted'])
def test_reference_test_and_example_scenes_match_oracle(name, gpu_engine, oracle, sims):
  '''
  scenes of the reference's own tests / example (test/50-old-tests grating.FCStd, gaussian.FCStd; examples/1-getting-started):
  gratings (wavelength plumbing, lineGrating), a mirror + plano-convex lens + absorber train
  '''
  sim = sims(name)
  n = 30000
  cfg = sim.cfg(record_all_hits=True, hit_capacity=12*n)
  sa = sim.source_args(0)
  with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sa), cfg, SEED, 0, n) as res:
    gc, gh = res.counts, res.hits(sort=True)
  o = oracle.trace_mc(sim.scene, sa, cfg, SEED, 0, n, hit_capacity=12*n, threads=0)
  assert compare_hits(gh, o['hits'], n) == 0
  assert gc == o['counts'] and gc['hits'] > n//2


@pytest.mark.parametrize('name', ['minimal', 'lensesAndMirrors', 'lensesAndMirrorsSequential'])
def test_recorded_hits_only(name, gpu_engine, oracle, sims):
  'default RecordHits flags: what the reference would store (absorber hits only)'
  sim = sims(name)
  n = 50000
  cfg = sim.cfg(hit_capacity=2*n)
  sa = sim.source_args(0)
  with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sa), cfg, SEED, 0, n) as res:
    gc, gh = res.counts, res.hits(sort=True)
  o = oracle.trace_mc(sim.scene, sa, cfg, SEED, 0, n, hit_capacity=2*n, threads=0)
  assert gc == o['counts']
  compare_hits(gh, o['hits'], n)
  assert set(gh['group']) == {len(sim.scene.groups)-1}


def explicit_fan(n=2000, spread=0.02, seed=5):
  rng = np.random.default_rng(seed)
  th, ph = np.abs(rng.normal(0, spread, n)), rng.uniform(0, 2*np.pi, n)
  d = np.stack([np.sin(th)*np.sin(ph), -np.sin(th)*np.cos(ph), np.cos(th)], axis=-1)
  return np.zeros_like(d), d


@pytest.mark.parametrize('name', ['minimal', 'lensesAndMirrors', 'lensesAndMirrorsSequential'])
def test_explicit_ray_list_matches_oracle(name, gpu_engine, oracle, sims):
  'odw_trace_rays (fans / replay / parity path): hits, per-ray segment count, final point and power'
  sim = sims(name)
  o_, d_ = explicit_fan(spread=0.5 if name == 'minimal' else 0.05)          # wide enough that some rays miss the optics
  d_ = d_*np.linspace(0.5, 2.0, len(d_))[:, None]    # directions need not be unit (traceRay normalises where needed)
  p_ = np.linspace(0.1, 1.0, len(d_))
  cfg = sim.cfg(record_all_hits=True, hit_capacity=20*len(d_))
  with gpu_engine.scene(sim.scene).trace_rays(cfg, o_, d_, p_) as res:
    gc, gh, gs = res.counts, res.hits(sort=True), res.ray_summary()
  o = oracle.trace_rays(sim.scene, cfg, o_, d_, p_, hit_capacity=20*len(d_))
  assert gc == o['counts']
  compare_hits(gh, o['hits'], len(d_))
  assert np.array_equal(gs['n_segments'], o['n_segments'])
  assert np.abs(gs['final_points']-o['final_points']).max() < 1e-6     # escape segments: 460 mm x direction error
  assert np.abs(gs['final_powers']-o['final_powers']).max() < 1e-14
  assert gc['escaped'] > 0


def procedural_scene():
  b = SceneBuilder()
  lens = b.add_group('Lens', 'Lens', optical_type='Lens', refractive_index=1.6)
  b.add_shape(lens, prim.plano_convex_lens(30.0, 6.0, 0.8), prim.translation(0, 0, 20) @ prim.rotation((1, 0, 0), 0.05))
  b.add_shape(lens, prim.cone(4.0, 2.0, 5.0), prim.translation(12, 0, 18))
  mir = b.add_group('Mirror', 'Mirror', optical_type='Mirror', reflectivity=0.9)
  b.add_shape(mir, prim.cylinder(3.0, 8.0), prim.translation(-12, 0, 30) @ prim.rotation((0, 1, 0), 1.2))
  b.add_shape(mir, prim.torus(8.0, 1.5), prim.translation(0, 0, 45))
  vac = b.add_group('Vac', 'Vac', optical_type='Vacuum', record_hits=True)
  b.add_shape(vac, prim.box(40, 40, 2), prim.translation(-20, -20, 55))
  gr = b.add_group('Grating', 'Grating', optical_type='Grating', grating_type='Transmission', refractive_index=1.5,
                   grating_lines_per_mm=300, grating_order=1, grating_orientation=(0, 1, 0))
  b.add_shape(gr, prim.box(60, 60, 1), prim.translation(-30, -30, 62))
  ab = b.add_group('Abs', 'Abs', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.disc(60.0), prim.translation(0, 0, 80))
  b.add_shape(ab, prim.sphere(2.0), prim.translation(0, 9, 35))
  return b.build()


def test_procedural_scene_all_surface_and_optical_types(gpu_engine, oracle):
  'cone, cylinder, torus, sphere cap, disc (arc loop), mirror/lens/vacuum/transmission grating/absorber, TIR'
  sc = procedural_scene()
  o_, d_ = explicit_fan(n=20000, spread=0.25, seed=9)
  cfg = _abi.CfgArgs(max_ray_length=300.0, dist_tol=1e-6, max_intersections=50, record_all_hits=True, hit_capacity=60*len(d_))
  with gpu_engine.scene(sc).trace_rays(cfg, o_, d_) as res:
    gc, gh, gs = res.counts, res.hits(sort=True), res.ray_summary()
  o = oracle.trace_rays(sc, cfg, o_, d_, wavelength=500.0, hit_capacity=60*len(d_), threads=0)
  bad = compare_hits(gh, o['hits'], len(d_), max_bad_fraction=1e-3)
  assert bad <= 20
  kinds = {int(sc.faces[f]['kind']) for f in np.unique(gh['face_id'])}
  assert kinds == {1, 2, 3, 4, 5}
  assert set(np.unique(gh['group'])) == {0, 1, 2, 3, 4}


def test_bvh_path_equals_brute_force_semantics(gpu_engine, oracle):
  'a >64-face scene takes the BVH kernel; the oracle (exhaustive loops) is the arbiter'
  rng = np.random.default_rng(2)
  b = SceneBuilder()
  groups = [b.add_group('L', 'L', optical_type='Lens', refractive_index=1.4),
            b.add_group('M', 'M', optical_type='Mirror'),
            b.add_group('A', 'A', optical_type='Absorber', record_hits=True)]
  for i in range(40):
    c = rng.uniform(-15, 15, 3) + [0, 0, 40]
    if i % 2:
      b.add_shape(groups[i % 3], prim.box(*rng.uniform(1, 4, 3)), prim.translation(*c) @ prim.rotation(rng.normal(size=3), rng.uniform(0, 3)))
    else:
      b.add_shape(groups[i % 3], prim.sphere(rng.uniform(0.5, 2.5)), prim.translation(*c))
  sc = b.build()
  assert len(sc.faces) > 64
  o_, d_ = explicit_fan(n=30000, spread=0.3, seed=4)
  cfg = _abi.CfgArgs(max_ray_length=200.0, max_intersections=30, record_all_hits=True, hit_capacity=40*len(d_))
  with gpu_engine.scene(sc).trace_rays(cfg, o_, d_) as res:
    gh = res.hits(sort=True)
  o = oracle.trace_rays(sc, cfg, o_, d_, hit_capacity=40*len(d_), threads=0)
  assert compare_hits(gh, o['hits'], len(d_), max_bad_fraction=1e-3, **CHAOTIC) <= 30


def test_edge_cases_empty_overflow_and_errors(gpu_engine, sims):
  sim = sims('lensesAndMirrors')
  ds = gpu_engine.scene(sim.scene)
  dsrc = gpu_engine.source(sim.source_args(0))
  # empty inputs
  with ds.trace_mc(dsrc, sim.cfg(), SEED, 0, 0) as res:
    assert res.counts['rays'] == 0 and res.counts['segments'] == 0 and len(res.hits()['powers']) == 0
  with ds.trace_rays(sim.cfg(), np.zeros((0, 3)), np.zeros((0, 3))) as res:
    assert res.counts['segments'] == 0
  # hit buffer too small: complete counters, ODW_EOVERFLOW reported, stored prefix valid
  with ds.trace_mc(dsrc, sim.cfg(hit_capacity=100), SEED, 0, 5000) as res:
    assert res.overflow and res.counts['hits_dropped'] > 0
    assert res.counts['hits'] == res.counts['hits_dropped']+100
    assert len(res.hits()['powers']) == 100
  # malformed descriptions fail loudly
  bad = sims('minimal').scene
  import copy
  sc2 = copy.copy(bad)
  sc2.faces = bad.faces.copy()
  sc2.faces['group'][0] = 7
  with pytest.raises(engine.EngineError):
    gpu_engine.scene(sc2)
  with pytest.raises(engine.EngineError):
    engine.Engine(99)


def test_kernel_instances_agree(gpu_engine, sims):
  '''
  The Monte-Carlo kernel exists in several feature sets (FEAT_* in csrc/odw_trace.cuh: none / sequential / all) and the
  host picks the leanest that covers a launch.  Same rays through the lean and the full instance (forced by asking for a
  device histogram, FEAT_BIN) must give the same hit list bit for bit; likewise sequential scene vs the same scene
  with a histogram.  Also covers dynamic ray claiming: one request, different launch sizes and stream counts.
  '''
  n = 300000
  ref = {}
  for name in ('lensesAndMirrors', 'lensesAndMirrorsSequential'):
    sim = sims(name)
    absorber = len(sim.scene.groups)-1
    spec = dict(group=absorber, nu=8, nv=8, origin=(-68.86, 0, 73), uaxis=(1, 0, 0), vaxis=(0, 1, 0), u_range=(-1, 1), v_range=(-1, 1))
    ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
    with ds.trace_mc(dsrc, sim.cfg(record_all_hits=True, hit_capacity=9*n), SEED, 123, n) as res:
      lean, c0 = res.hits(sort=True), res.counts
    with ds.trace_mc(dsrc, sim.cfg(record_all_hits=True, hit_capacity=9*n, binnings=[spec]), SEED, 123, n) as res:
      full, c1 = res.hits(sort=True), res.counts
    assert c0['segments'] == c1['segments'] and c0['hits'] == c1['hits'] > 6*n
    for key in lean:
      assert np.array_equal(lean[key], full[key]), (name, key)
    ref[name] = full
  sim = sims('lensesAndMirrors')
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  for wave, streams in (('4096', '1'), ('65536', '3'), ('1000003', '8')):
    os.environ['ODW_RAYS_PER_LAUNCH'], os.environ['ODW_STREAMS'] = wave, streams
    try:
      with ds.trace_mc(dsrc, sim.cfg(record_all_hits=True, hit_capacity=9*n), SEED, 123, n) as res:
        h = res.hits(sort=True)
    finally:
      del os.environ['ODW_RAYS_PER_LAUNCH'], os.environ['ODW_STREAMS']
    for key in h:
      assert np.array_equal(h[key], ref['lensesAndMirrors'][key]), (wave, key)


def test_count_only_and_histogram_modes(gpu_engine, oracle, sims):
  'device binning == numpy.histogram2d of the stored hit list (reference jupyter_utils/histogram.py:54 semantics)'
  sim = sims('lensesAndMirrors')
  n = 200000
  absorber = len(sim.scene.groups)-1
  spec = dict(group=absorber, nu=40, nv=30, origin=(-68.86, 0, 73), uaxis=(1, 0, 0), vaxis=(0, 1, 0),
              u_range=(-0.4, 0.4), v_range=(-0.3, 0.3))
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  with ds.trace_mc(dsrc, sim.cfg(binnings=[spec], hit_capacity=2*n), SEED, 0, n) as res:
    hist, hits, c1 = res.histogram(0), res.hits(sort=False), res.counts
  x, y = hits['points'][:, 0]+68.86, hits['points'][:, 1]
  ref, _, _ = np.histogram2d(x, y, bins=(40, 30), range=((-0.4, 0.4), (-0.3, 0.3)))
  assert hist.sum() > 0.5*n
  assert np.abs(hist-ref).sum() <= 2            # a hit exactly on a bin edge may fall either side
  with ds.trace_mc(dsrc, sim.cfg(binnings=[spec], store_hits=False), SEED, 0, n) as res:
    assert np.array_equal(res.histogram(0), hist) and res.counts['hits'] == c1['hits']
    assert res.counts['segments'] == c1['segments']
  o = oracle.trace_mc(sim.scene, sim.source_args(0), sim.cfg(binnings=[spec], store_hits=False), SEED, 0, n, threads=0)
  assert np.abs(o['histograms'][0]-hist).sum() <= 2


def test_host_delivery_equals_device_hit_list(gpu_engine, sims):
  'odw_trace_mc_host (chunked, copy overlapped with compute) delivers exactly the rows of odw_trace_mc'
  sim = sims('lensesAndMirrors')
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  n = 300000
  os.environ['ODW_HOST_CHUNK'] = '70000'       # several chunks, last one ragged
  try:
    arrays = _abi.HitArrays(n+16)
    counts, got = ds.trace_mc_host(dsrc, sim.cfg(), SEED, 123, n, arrays.view)
  finally:
    del os.environ['ODW_HOST_CHUNK']
  host = arrays.trimmed(got, sort=True)
  with ds.trace_mc(dsrc, sim.cfg(hit_capacity=2*n), SEED, 123, n) as res:
    dev, c = res.hits(sort=True), res.counts
  assert got == c['hits'] and counts['segments'] == c['segments'] and counts['escaped'] == c['escaped']
  for key in dev:
    assert np.array_equal(host[key], dev[key]), key
  # too small a host buffer: reported, prefix still valid
  small = _abi.HitArrays(1000)
  counts, got = ds.trace_mc_host(dsrc, sim.cfg(), SEED, 123, n, small.view)
  assert got == 1000 and counts['hits_dropped'] == c['hits']-1000


def test_host_delivery_many_waves_on_many_streams(gpu_engine, sims, monkeypatch):
  """
  Several chunks of several launch waves rotating over 4 streams: the per-wave claim counters are zeroed on the engine
  stream BEFORE the fork, so no wave stream can read the stale counters of the previous chunk (which would skip or
  repeat a whole wave).  Repeated, because the failure was a race.
  """
  sim = sims('lensesAndMirrors')
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  n = 300000
  with ds.trace_mc(dsrc, sim.cfg(hit_capacity=2*n), SEED, 7, n) as res:
    dev, c = res.hits(sort=True), res.counts
  monkeypatch.setenv('ODW_HOST_CHUNK', '65536')
  monkeypatch.setenv('ODW_RAYS_PER_LAUNCH', '8192')
  monkeypatch.setenv('ODW_STREAMS', '4')
  arrays = _abi.HitArrays(n+16)
  for _ in range(5):
    counts, got = ds.trace_mc_host(dsrc, sim.cfg(), SEED, 7, n, arrays.view)
    assert counts['waves'] >= 4*8 + 4
    assert got == c['hits'] and counts['segments'] == c['segments'] and counts['escaped'] == c['escaped']
    host = arrays.trimmed(got, sort=True)
    for key in dev:
      assert np.array_equal(host[key], dev[key]), key


def test_host_delivery_wavefront_chunks_keep_their_counters(gpu_engine, sims, monkeypatch):
  'BVH / wavefront scenes: the per-bounce survivor count has its own pinned word, the chunk counters stay intact'
  monkeypatch.setenv('ODW_BVH', '1')
  sim = sims('lensesAndMirrors')
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  n = 200000
  with ds.trace_mc(dsrc, sim.cfg(hit_capacity=2*n), SEED, 0, n) as res:
    dev, c = res.hits(sort=True), res.counts
  monkeypatch.setenv('ODW_HOST_CHUNK', '50000')
  arrays = _abi.HitArrays(n+16)
  counts, got = ds.trace_mc_host(dsrc, sim.cfg(), SEED, 0, n, arrays.view)
  assert counts['segments'] == c['segments'] and got == c['hits'] and counts['escaped'] == c['escaped']
  host = arrays.trimmed(got, sort=True)
  for key in dev:
    assert np.array_equal(host[key], dev[key]), key


def test_host_delivery_more_than_two_hits_per_ray(gpu_engine, sims, monkeypatch):
  'hit_capacity sizes the per-chunk device lists too: a scene recording ~7 hits per ray is delivered complete'
  sim = sims('lensesAndMirrors')
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  n = 100000
  with ds.trace_mc(dsrc, sim.cfg(hit_capacity=8*n, record_all_hits=True), SEED, 0, n) as res:
    dev, c = res.hits(sort=True), res.counts
  assert c['hits'] > 6*n and c['hits_dropped'] == 0
  monkeypatch.setenv('ODW_HOST_CHUNK', '30000')
  arrays = _abi.HitArrays(8*n)
  counts, got = ds.trace_mc_host(dsrc, sim.cfg(record_all_hits=True), SEED, 0, n, arrays.view)       # default: two rows per ray
  assert counts['hits_dropped'] > 0 and counts['hits'] == c['hits']
  counts, got = ds.trace_mc_host(dsrc, sim.cfg(record_all_hits=True, hit_capacity=8*n), SEED, 0, n, arrays.view)
  assert counts['hits_dropped'] == 0 and got == c['hits']
  host = arrays.trimmed(got, sort=True)
  for key in dev:
    assert np.array_equal(host[key], dev[key]), key


def test_convex_shell_skip_changes_only_edge_rays(gpu_engine, sims, monkeypatch):
  """
  The kernel skips a convex shell for the segment that starts on it and points away from it (odw_trace.cuh interact());
  the reference tests the shell and finds nothing beyond distTol — unless the segment starts in the tolerance zone just
  OUTSIDE the solid (a hit accepted up to distTol beyond a face edge), where it can meet a neighbouring face of the same
  shell again.  ODW_SKIP_CONVEX=0 switches the shortcut off.  With and without it every hit row must be bit-equal except
  for rays that have a hit within 2 distTol of a face edge (checked with the stand-in geometry of tests/occ_stub.py); on
  the headline scene (distTol = 1e-6 mm) no ray of 2e5 may differ at all.
  """
  import traceray_cases as cases
  from test_traceray_golden import edge_distance
  builds = (cases.glass_ball, cases.curved_surfaces, cases.slab_stack, cases.behind_the_start, cases.placed_groups)

  def run_all():
    out = []
    for build in builds:
      scene, o, d, settings = build()
      cfg = cases.synthetic_cfg(settings, record_all_hits=True, hit_capacity=60*len(o))
      ds = gpu_engine.scene(scene)
      with ds.trace_rays(cfg, o, d) as res:
        out.append((res.hits(sort=True), res.ray_summary(), scene, len(o)))
      ds.close()
    sim = sims('lensesAndMirrors')
    ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
    with ds.trace_mc(dsrc, sim.cfg(record_all_hits=True, hit_capacity=2_000_000), SEED, 0, 200000) as res:
      out.append((res.hits(sort=True), dict(counts=np.array(list(res.counts.values()))), None, 200000))
    ds.close(); dsrc.close()
    return out

  monkeypatch.setenv('ODW_SKIP_CONVEX', '1')
  with_skip = run_all()
  monkeypatch.setenv('ODW_SKIP_CONVEX', '0')
  without = run_all()
  n_edge = 0
  for (ha, sa, scene, n), (hb, sb, _, _) in zip(with_skip, without):
    same = len(ha['face_id']) == len(hb['face_id']) and all(np.array_equal(ha[k], hb[k]) for k in ha)
    if same:
      for key in sa:
        assert np.array_equal(sa[key], sb[key]), key
      continue
    assert scene is not None, 'the headline scene must not depend on the convex-shell skip'
    seq_a, seq_b = per_ray_sequences(ha, n), per_ray_sequences(hb, n)
    bad = [r for r in range(n) if seq_a[r] != seq_b[r]]
    assert 0 < len(bad) <= 0.01*n
    rows = {int(f['face_id']): i for i, f in enumerate(scene.faces)}
    for r in bad:
      sel = hb['ray_index'] == r
      near = min(edge_distance(scene, rows[int(f)], P) for f, P in zip(hb['face_id'][sel], hb['points'][sel]))
      assert near < 2*cases.TOL, f'ray {r} depends on the convex-shell skip without being an edge ray ({near:.3g})'
    n_edge += len(bad)
    keep_a, keep_b = ~np.isin(ha['ray_index'], bad), ~np.isin(hb['ray_index'], bad)
    for key in ha:
      assert np.array_equal(ha[key][keep_a], hb[key][keep_b]), key
  assert n_edge > 0          # curved_surfaces has such rays (the rim of the frustum): the exemption is exercised, not vacuous


@pytest.mark.parametrize('name,n', [('lensesAndMirrors', 1 << 22), ('hugeArray', 1 << 21)])
def test_append_and_compaction_invariants(name, n, gpu_engine, sims):
  """
  Stand-in for the racecheck / memcheck runs this pool does not allow (profiles/sanitizer_r02.txt): size-independent
  invariants of the warp-aggregated hit append (register-resident kernel, waves on 4 streams) and of the ballot
  compaction of the wavefront kernels.  With every intersection recorded, the rows of a ray are exactly its bounces
  0..k-1, no (ray, bounce) pair occurs twice, rows == segments - escaped, and a second run gives the same sorted rows.
  """
  sim = sims(name)
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  cap = n*(9 if name == 'lensesAndMirrors' else 40)
  runs = []
  for _ in range(2):
    with ds.trace_mc(dsrc, sim.cfg(record_all_hits=True, hit_capacity=cap), SEED, 1 << 33, n) as res:
      c = res.counts
      arrays = _abi.HitArrays(c['hits'])
      for k in ('directions', 'powers', 'is_entering', 'group', 'medium'):      # not needed here: skip their copies
        setattr(arrays.view, k, None)
      got = C.c_uint64(0)
      engine._check(engine.load_library().odw_result_hits(res._h, C.addressof(arrays.view), 0, C.byref(got)))
    assert c['hits_dropped'] == 0 and got.value == c['hits'] == c['segments']-c['escaped']
    ray = arrays.ray_index[:got.value].astype(np.int64) - (1 << 33)
    bounce = arrays.bounce[:got.value].astype(np.int64)
    assert ray.min() >= 0 and ray.max() < n and bounce.min() == 0 and bounce.max() < 128
    order = np.argsort(ray*128 + bounce, kind='stable')
    ray, bounce = ray[order], bounce[order]
    start = np.searchsorted(ray, np.arange(n))
    assert np.array_equal(bounce, np.arange(len(ray)) - start[ray])            # each ray: bounces 0..k-1, each once
    runs.append((ray, bounce, arrays.face_id[:got.value][order].copy(), arrays.points[:got.value][order].copy(), c))
  (r0, b0, f0, p0, c0), (r1, b1, f1, p1, c1) = runs
  assert c0 == c1 and np.array_equal(r0, r1) and np.array_equal(f0, f1) and np.array_equal(p0, p1)
  ds.close(); dsrc.close()


def test_range_splitting_is_invariant(gpu_engine, sims):
  'Philox counter = global ray index: tracing [0,n) equals tracing [0,k) and [k,n) (GPU-count invariance, SURVEY §8e)'
  sim = sims('lensesAndMirrors')
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  n, k = 100000, 37777
  cfg = sim.cfg(hit_capacity=2*n)
  with ds.trace_mc(dsrc, cfg, SEED, 0, n) as res:
    whole, cw = res.hits(sort=True), res.counts
  parts, segs = [], 0
  for first, cnt in ((0, k), (k, n-k)):
    with ds.trace_mc(dsrc, cfg, SEED, first, cnt) as res:
      parts.append(res.hits(sort=True)); segs += res.counts['segments']
  assert segs == cw['segments']
  for key in whole:
    assert np.array_equal(np.concatenate([p[key] for p in parts]), whole[key]), key


def test_full_size_properties(gpu_engine, sims):
  '''
  BASELINE.json configs[1] size (1e8 rays, lensesAndMirrors): size-independent properties instead of an
  oracle run — counter identities, every stored hit lies on the absorber's entry faces, power 1, unique rays.
  '''
  sim = sims('lensesAndMirrors')
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  n = int(os.environ.get('ODW_FULL_RAYS', '100000000'))
  cfg = sim.cfg(hit_capacity=n+1024)
  with ds.trace_mc(dsrc, cfg, SEED, 0, n) as res:
    c = res.counts
    assert c['hits_dropped'] == 0 and c['rays'] == n and c['depth_terminated'] == 0
    # each ray either ends on the absorber (one recorded hit) or leaves the scene (escape segment)
    assert c['hits']+c['escaped'] == n
    assert 6.99*n < c['segments'] < 7.01*n        # a few rays take extra internal reflections at the lens rims
    arrays = _abi.HitArrays(c['hits'])
    h = res.hits(sort=False, into=arrays)
  assert len(h['powers']) == c['hits']
  assert np.all(h['powers'] == 1.0) and np.all(h['is_entering'] == 1) and np.all(h['group'] == 3)
  z = h['points'][:, 2]
  on_box = np.abs(z-73.0) < 1e-9
  assert on_box.mean() > 0.999                        # the rest end on the torus
  assert np.unique(h['ray_index']).size == c['hits']  # one recorded hit per ray at most
  # beam centre and symmetric spread on the detector (on-axis answer of SURVEY.md Appendix B)
  assert abs(h['points'][on_box, 0].mean()+68.857864) < 1e-3 and abs(h['points'][on_box, 1].mean()) < 1e-3


@pytest.mark.parametrize('tail', ['0', '100000000'])
def test_wavefront_kernels_equal_register_resident_kernel(gpu_engine, oracle, sims, monkeypatch, tail):
  '''
  The two formulations of the trace (odw_kernels.cu: one lane owns a ray for life; odw_wavefront.cu: generate /
  traverse / interact per bounce through a ray pool in HBM) must produce the same hits.  ODW_BVH=1 sends a small scene
  down the BVH path; tail = 0: every bounce through traverse + interact, tail = huge: everything in the tail kernel.
  '''
  sim = sims('lensesAndMirrors')
  sa = sim.source_args(0)
  n = 40000
  cfg = sim.cfg(record_all_hits=True, hit_capacity=10*n)
  with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sa), cfg, SEED, 5000, n) as res:
    c0, h0 = res.counts, res.hits(sort=True)
  monkeypatch.setenv('ODW_BVH', '1')
  monkeypatch.setenv('ODW_WAVEFRONT', '1')
  monkeypatch.setenv('ODW_WAVEFRONT_TAIL', tail)
  ds = gpu_engine.scene(sim.scene)                       # the knobs are read when the scene is created / traced
  with ds.trace_mc(gpu_engine.source(sa), cfg, SEED, 5000, n) as res:
    c1, h1 = res.counts, res.hits(sort=True)
  o_, d_ = explicit_fan(n=5000, spread=0.05)
  d_ = d_*np.linspace(0.5, 2.0, len(d_))[:, None]
  cfg2 = sim.cfg(record_all_hits=True, hit_capacity=20*len(d_))
  with ds.trace_rays(cfg2, o_, d_) as res:
    c3, h3, s3 = res.counts, res.hits(sort=True), res.ray_summary()
  monkeypatch.delenv('ODW_BVH'); monkeypatch.delenv('ODW_WAVEFRONT'); monkeypatch.delenv('ODW_WAVEFRONT_TAIL')
  for k in ('rays', 'segments', 'hits', 'escaped', 'depth_terminated'):
    assert c0[k] == c1[k], k
  assert compare_hits(h1, h0, n, base=5000) == 0
  o = oracle.trace_rays(sim.scene, cfg2, o_, d_, hit_capacity=20*len(d_))
  assert compare_hits(h3, o['hits'], len(d_)) == 0
  assert np.array_equal(s3['n_segments'], o['n_segments'])
  assert np.abs(s3['final_points']-o['final_points']).max() < 1e-6
