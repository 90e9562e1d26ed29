'''
Surface light source (reference freecad_elements/surface_source.py): host-side face areas / weights, the oracle's
sampler (area-uniform positions on trimmed faces, theta without sin(theta), phi uniform, direction formula of
_makeRay) and — on the GPU — bit-for-bit the same rays and hits from the CUDA kernel.

Pinned against the reference: the theta table (reference ScalarRandomVariable, tests/golden/sampler_golden.npz
covers the same table code) and the scene/emitter selection of test/21-simulation-modes/main.FCStd.  The placement of
points on faces is "parity unpinned" (the reference asks OpenCASCADE for valueAt / distToShape); it is checked
against analytic area measures here.
'''
import os

import numpy as np
import pytest
from scipy import stats

from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.distributions import surface_source_tables
from freecad.optics_design_workbench_b200.freecad_elements import surface_source
from freecad.optics_design_workbench_b200.scene_export import primitives as prim, scene as sc
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder

from conftest import SCENES, SEED


def emitter(shapes):
  'shapes: list of (FaceInstance list, transform) -> EmittingFaces'
  return surface_source.emitting_faces_from_instances(shapes)


def source_args(emit, density='cos(theta)**2', domain='0, pi/2', res='1e4', source_id=0):
  rec = dict(PowerDensity=density, ThetaDomain=domain, ThetaResolutionNumericMode=res)
  return _abi.SourceArgs(surface_source_tables(rec), kind=sc.SRC_SURFACE, source_id=source_id, gpM=np.eye(4),
                         emit_faces=emit.faces, emit_segs=emit.segs, emit_cdf=emit.cdf, dist_tol=1e-6)


def all_kinds_emitter():
  return emitter([
    (prim.box(4, 6, 2), prim.translation(-10, 0, 0)),                                  # 6 rectangles
    (prim.disc(3.0), prim.translation(0, 10, 0) @ prim.rotation((1, 0, 0), 0.7)),      # plane + arc loop
    (prim.cylinder(2.0, 5.0), prim.translation(10, 0, 0)),                             # wall + 2 discs
    (prim.sphere(2.5), prim.translation(0, -10, 0)),                                   # untrimmed sphere
    (prim.plano_convex_lens(8.0, 3.0, 0.5), prim.translation(0, 0, 12)),               # spherical cap + band + disc
    (prim.cone(3.0, 1.0, 4.0), prim.translation(15, 15, 0)),
    (prim.torus(6.0, 1.0), prim.translation(-15, -15, 0)),
  ])


def test_face_areas_closed_form_and_scanline():
  e = emitter([(prim.box(4, 6, 2), np.eye(4))])
  assert sorted(np.round(e.areas, 12)) == [8, 8, 12, 12, 24, 24]
  e = emitter([(prim.sphere(2.5), np.eye(4))])
  np.testing.assert_allclose(e.areas.sum(), 4*np.pi*2.5**2, rtol=1e-12)
  e = emitter([(prim.torus(6.0, 1.0), np.eye(4))])
  np.testing.assert_allclose(e.areas.sum(), 4*np.pi**2*6.0*1.0, rtol=1e-12)
  e = emitter([(prim.cylinder(2.0, 5.0), np.eye(4))])
  np.testing.assert_allclose(e.areas.sum(), 2*np.pi*2*5 + 2*np.pi*4, rtol=1e-12)
  e = emitter([(prim.cone(3.0, 1.0, 4.0), np.eye(4))])
  slant = np.hypot(4.0, 2.0)
  np.testing.assert_allclose(e.areas.sum(), np.pi*(3+1)*slant + np.pi*9 + np.pi*1, rtol=1e-9)
  # scanline integral of a loop-trimmed face against its closed form (the disc shortcut bypassed)
  d = emitter([(prim.disc(3.0), np.eye(4))])
  np.testing.assert_allclose(surface_source._scanline_area(d.faces[0], d.segs), np.pi*9, rtol=2e-4)
  c = emitter([(prim.cylinder(2.0, 5.0), np.eye(4))])
  wall = [f for f in c.faces if int(f['kind']) == sc.SURF_CYLINDER][0]
  assert abs(c.cdf[-1]-1) == 0 and np.all(np.diff(c.cdf) > 0)
  np.testing.assert_allclose(surface_source.face_area(wall, c.segs), 2*np.pi*2*5, rtol=1e-12)


def test_points_lie_on_the_faces_and_faces_are_chosen_by_area(oracle):
  e = all_kinds_emitter()
  sa = source_args(e)
  n = 400000
  s = oracle.sample_mc(sa, SEED, 0, n)
  P, D = s['origins'], s['directions']
  # which face does each point lie on?  distance to every emitting surface, trimmed by its window
  counts = np.zeros(len(e.faces))
  owner = np.full(n, -1)
  for k, f in enumerate(e.faces):
    w = P-f['origin']
    x, y, z = w@f['xdir'], w@f['ydir'], w@f['zdir']
    kind = int(f['kind'])
    if kind == sc.SURF_PLANE:
      dist = np.abs(z)
      inside = np.ones(n, bool)
      if int(f['trim_kind']) == sc.TRIM_UVBOX:
        inside = (x >= f['uv_min'][0]-1e-9) & (x <= f['uv_max'][0]+1e-9) & (y >= f['uv_min'][1]-1e-9) & (y <= f['uv_max'][1]+1e-9)
      else:
        seg = e.segs[int(f['seg_first'])]
        inside = np.hypot(x-seg['a'][0], y-seg['a'][1]) <= seg['a'][2]+1e-9
    elif kind == sc.SURF_CYLINDER:
      dist, inside = np.abs(np.hypot(x, y)-f['p0']), (z >= f['uv_min'][1]-1e-9) & (z <= f['uv_max'][1]+1e-9)
    elif kind == sc.SURF_SPHERE:
      dist = np.abs(np.sqrt(x*x+y*y+z*z)-f['p0'])
      v = np.arcsin(np.clip(z/f['p0'], -1, 1))
      inside = np.ones(n, bool) if int(f['trim_kind']) == sc.TRIM_NONE else (v >= f['uv_min'][1]-1e-9) & (v <= f['uv_max'][1]+1e-9)
    elif kind == sc.SURF_CONE:
      v = z/np.cos(f['p1'])
      dist, inside = np.abs(np.hypot(x, y)-np.abs(f['p0']+v*np.sin(f['p1']))), (v >= f['uv_min'][1]-1e-9) & (v <= f['uv_max'][1]+1e-9)
    else:
      dist, inside = np.abs(np.hypot(np.hypot(x, y)-f['p0'], z)-f['p1']), np.ones(n, bool)
    mine = (dist < 1e-9) & inside & (owner < 0)
    owner[mine] = k
    counts[k] = mine.sum()
  assert (owner >= 0).all()                                            # every origin is on an emitting face
  expected = e.areas/e.areas.sum()*n
  chi2 = ((counts-expected)**2/expected).sum()
  assert stats.chi2.sf(chi2, len(counts)-1) > 1e-3, (counts, expected)  # face choice ~ area (surface_source.py:465-466)
  # direction: theta measured from the OUTWARD normal; unit length
  np.testing.assert_allclose(np.linalg.norm(D, axis=1), 1.0, atol=1e-12)
  assert s['first'].min() >= 0 and s['first'].max() <= np.pi/2 and s['phi'].min() >= 0 and s['phi'].max() < 2*np.pi


def test_positions_are_uniform_by_area_on_curved_faces(oracle):
  'sphere: z uniform (Archimedes); torus: density (R + r cos v); cone: density ~ radius; disc: r^2 uniform'
  n = 300000
  def ks(sample, cdf):
    return stats.kstest(sample, cdf).pvalue
  s = oracle.sample_mc(source_args(emitter([(prim.sphere(2.5), prim.translation(1, 2, 3))])), SEED, 0, n)
  z = (s['origins'][:, 2]-3)/2.5
  assert ks(z, stats.uniform(-1, 2).cdf) > 1e-3
  assert ks(np.arctan2(s['origins'][:, 1]-2, s['origins'][:, 0]-1), stats.uniform(-np.pi, 2*np.pi).cdf) > 1e-3
  R, r = 6.0, 1.0
  s = oracle.sample_mc(source_args(emitter([(prim.torus(R, r), np.eye(4))])), SEED, 0, n)
  P = s['origins']
  v = np.arctan2(P[:, 2], np.hypot(P[:, 0], P[:, 1])-R) % (2*np.pi)
  assert ks(v, lambda x: (R*x + r*np.sin(x))/(2*np.pi*R)) > 1e-3
  e = emitter([(prim.cone(3.0, 1.0, 4.0), np.eye(4))])
  cone_only = surface_source.EmittingFaces(e.faces[[int(f['kind']) == sc.SURF_CONE for f in e.faces]], e.segs)
  s = oracle.sample_mc(source_args(cone_only), SEED, 0, n)
  zc = s['origins'][:, 2]                                            # radius falls linearly 3 -> 1 over z in [0, 4]
  assert ks(zc, lambda x: (3*x - x*x/4)/8) > 1e-3
  s = oracle.sample_mc(source_args(emitter([(prim.disc(3.0), np.eye(4))])), SEED, 0, n)
  rr = (s['origins'][:, 0]**2+s['origins'][:, 1]**2)/9
  assert rr.max() <= 1+1e-6 and ks(rr, stats.uniform(0, 1).cdf) > 1e-3  # rejection inside the arc loop, dilated by dist_tol like distToShape < tol


def test_direction_formula_and_angular_density(oracle):
  'd = cos(theta) n + sin(theta) (cos(phi) (t x n) + sin(phi) t), theta ~ cos^2 WITHOUT sin(theta) factor, phi uniform'
  e = emitter([(prim.rectangle(10, 10), np.eye(4))])
  f = e.faces[0]
  sa = source_args(e)
  n = 300000
  s = oracle.sample_mc(sa, SEED, 0, n)
  nrm = f['zdir']*f['nsign']
  t = f['xdir']
  txn = np.cross(t, nrm)
  th, ph = s['first'], s['phi']
  want = (np.cos(th)[:, None]*nrm + np.sin(th)[:, None]*(np.cos(ph)[:, None]*txn + np.sin(ph)[:, None]*t))
  np.testing.assert_allclose(s['directions'], want, atol=1e-12)
  # cos^2(theta) on [0, pi/2]: CDF (theta + sin(2 theta)/2)/(pi/2)
  assert stats.kstest(th, lambda x: (x + np.sin(2*x)/2)/(np.pi/2)).pvalue > 1e-3
  assert stats.kstest(ph, stats.uniform(0, 2*np.pi).cdf).pvalue > 1e-3
  assert abs(np.corrcoef(th, ph)[0, 1]) < 0.01


def test_reference_test21_scene_emits_from_box_face5(oracle, sims):
  'test/21-simulation-modes/main.FCStd: ActiveSurfaces = Box001.Face5 -> the z = 38 face of the 10x10 box, pointing to -z'
  sim = sims('surfaceSourceTest21')
  rec = sim.source_records[0]
  assert rec['proxy'] == 'SurfaceSourceProxy' and rec['PowerDensity'] == 'cos(theta)**2'
  sa = sim.source_args(0)
  assert sa.desc.kind == sc.SRC_SURFACE and sa.desc.n_emit == 1
  np.testing.assert_allclose(rec['emit'].areas, [100.0])
  s = oracle.sample_mc(sa, SEED, 0, 50000)
  np.testing.assert_allclose(s['origins'][:, 2], 38.0, atol=1e-12)
  assert np.abs(s['origins'][:, :2]).max() <= 5+1e-9 and s['directions'][:, 2].max() < 0
  assert s['first'].max() <= np.pi/4+1e-12
  # the reference's own assertions for this scene (test/21 run-simulations.py:42-64): hits are recorded, rays end
  r = oracle.trace_mc(sim.scene, sa, sim.cfg(), SEED, 0, 20000, threads=0)
  assert r['counts']['hits'] > 999 and r['counts']['depth_terminated'] == 0


def test_simulation_loop_runs_a_surface_source(tmp_path, sims):
  from freecad.optics_design_workbench_b200.simulation import simulation_loop
  from oracle_engine import OracleEngine
  from test_simulation_loop import load_hits
  sim = sims('surfaceSourceTest21')
  run = simulation_loop.runSimulation(sim, 'true', engine=OracleEngine(), basePath=str(tmp_path/'s.OpticsDesign'),
                                      settings=dict(EndAfterHits=1e3), maxBatchRays=2000)
  assert len(load_hits(run)['points']) > 999


# ------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_gpu_surface_source_draws_equal_oracle(gpu_engine, oracle):
  sa = source_args(all_kinds_emitter())
  n = 200000
  g = gpu_engine.source(sa).sample(SEED, 1000, n)
  o = oracle.sample_mc(sa, SEED, 1000, n)
  for k in ('first', 'phi', 'origins', 'directions'):
    np.testing.assert_allclose(g[k], o[k], rtol=0, atol=1e-9, err_msg=k)


@pytest.mark.gpu
def test_gpu_surface_source_trace_equals_oracle(gpu_engine, oracle, sims):
  sim = sims('surfaceSourceTest21')
  sa = sim.source_args(0)
  n = 50000
  cfg = sim.cfg(record_all_hits=True, hit_capacity=8*n)
  with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sa), cfg, SEED, 0, n) as res:
    gc, gh = res.counts, res.hits(sort=True)
  o = oracle.trace_mc(sim.scene, sa, cfg, SEED, 0, n, hit_capacity=8*n, threads=0)
  drop = lambda c: {k: v for k, v in c.items() if k != 'waves'}      # kernel launches: not a property of the result
  assert drop(gc) == drop(o['counts'])
  oh = o['hits']
  assert np.array_equal(gh['ray_index'], oh['ray_index']) and np.array_equal(gh['face_id'], oh['face_id'])
  np.testing.assert_allclose(gh['points'], oh['points'], rtol=0, atol=1e-9)
  np.testing.assert_allclose(gh['directions'], oh['directions'], rtol=0, atol=1e-9)


def test_tessellated_emitter_triangles_are_sampled_uniformly(oracle, sims):
  '''
  lambert-source.FCStd (reference test/50-old-tests): a B-spline emitter meshed into 65 700 triangles + one sphere zone.
  Triangles are sampled directly (no rejection): every draw lies inside its triangle's plane and the share of draws per
  face follows the area weights.
  '''
  sim = sims('lambertSource')
  sa = sim.source_args(0)
  emit = sim.source_records[0]['emit']
  n = 400000
  s = oracle.sample_mc(sa, SEED, 0, n)
  w = np.diff(np.concatenate([[0.0], emit.cdf]))
  sphere = emit.faces['kind'] == 4
  centre, radius = emit.faces['origin'][sphere][0], emit.faces['p0'][sphere][0]
  on_sphere = np.abs(np.linalg.norm(s['origins']-centre, axis=1)-radius) < 1e-9
  share = np.count_nonzero(on_sphere)/n
  assert abs(share-w[sphere].sum()) < 5*np.sqrt(share*(1-share)/n)
  lo, hi = emit.faces['aabb_min'].min(axis=0), emit.faces['aabb_max'].max(axis=0)
  assert (s['origins'] >= lo-1e-9).all() and (s['origins'] <= hi+1e-9).all()
  assert np.abs(np.linalg.norm(s['directions'], axis=1)-1).max() < 1e-12


@pytest.mark.gpu
def test_gpu_tessellated_emitter_equals_oracle(gpu_engine, oracle, sims):
  'guide-table face pick + direct triangle sampling on the device = the oracle, draw for draw; and a traced batch'
  sim = sims('lambertSource')
  sa = sim.source_args(0)
  n = 200000
  g = gpu_engine.source(sa).sample(SEED, 77, n)
  o = oracle.sample_mc(sa, SEED, 77, n)
  for k in ('first', 'phi', 'origins', 'directions'):
    np.testing.assert_allclose(g[k], o[k], rtol=0, atol=1e-9, err_msg=k)
  n = 50000
  cfg = sim.cfg(record_all_hits=True, hit_capacity=8*n)
  with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sa), cfg, SEED, 0, n) as res:
    gc, gh = res.counts, res.hits(sort=True)
  want = oracle.trace_mc(sim.scene, sa, cfg, SEED, 0, n, hit_capacity=8*n, threads=0)
  # DistanceTolerance of this project is 1e-2 mm and its detector sheets are 0.1 mm thick: rays within tolerance of a sheet's
  # edge are exempt (north star); everything else must agree row for row
  from test_gpu_parity import compare_hits
  bad = compare_hits(gh, want['hits'], n, max_bad_fraction=1e-3)
  assert abs(gc['segments']-want['counts']['segments']) <= 4*max(bad, 1)


@pytest.mark.gpu
def test_gpu_presampled_wave_order_does_not_show(gpu_engine, sims, monkeypatch):
  '''
  The rays of a surface-source wave are drawn by sample_kernel and traced in coherence order (radix sort of ray_sort_key,
  csrc/odw_api.cu launch_waves).  Order of tracing and number of waves must not change one bit of the result (same counters, same hit rows once sorted by
  (ray, bounce)); the in-kernel draw (ODW_NO_PRESAMPLE) gives the same rows to the last bit but one.
  '''
  sim = sims('lambertSource')
  sa = sim.source_args(0)
  n = 300000
  cfg = sim.cfg(record_all_hits=True, hit_capacity=8*n)
  def run():
    with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sa), cfg, SEED, 12345, n) as res:
      c = {k: v for k, v in res.counts.items() if k not in ('waves', 'sm_clock_khz')}
      return c, res.hits(sort=True)
  base_c, base_h = run()
  for env in ({'ODW_PRESAMPLE_SORT': '0'}, {'ODW_PRESAMPLE_SORT': '0,32'}, {'ODW_RAYS_PER_LAUNCH': '65536'}, {'ODW_NO_PRESAMPLE': '1'}):
    with monkeypatch.context() as m:
      for k, v in env.items():
        m.setenv(k, v)
      c, h = run()
    assert c == base_c, env
    for k in base_h:
      if 'ODW_NO_PRESAMPLE' in env and h[k].dtype.kind == 'f':
        # the list kernel normalises the stored unit direction once more: a last-bit difference, nothing else
        np.testing.assert_allclose(h[k], base_h[k], rtol=0, atol=1e-12, err_msg=k)
      else:
        assert np.array_equal(h[k], base_h[k]), (env, k)
