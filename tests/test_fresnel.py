'''
Fresnel opt-in (include/odw.h odw_group.fresnel): the north star names "Snell, Fresnel, TIR"; the reference has no
Fresnel split (ray.py:165-211, 488-495 refract every ray at a Lens face without loss), so the switch defaults OFF and
everything else in the repo is bit-identical to before.  With it ON a ray is reflected at a Lens face with the unpolarised
Fresnel reflectance and refracted otherwise (one ray in, one ray out; the choice is driven by the ray's Philox stream).
Known answers: normal incidence R = ((n-1)/(n+1))^2, R -> 1 towards grazing incidence, Brewster's angle for the p part,
energy conservation (every ray is either reflected or transmitted), TIR untouched.  CPU: the oracle.  GPU: the CUDA path
through the C ABI, row for row against the oracle.
'''
import numpy as np
import pytest

from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.scene_export import primitives as prim
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder

N_INDEX = 1.5


def slab_scene(fresnel, n=N_INDEX):
  'a thick glass slab (z in [10, 20]) inside an absorbing sphere'
  b = SceneBuilder()
  lens = b.add_group('Slab', 'Slab', optical_type='Lens', refractive_index=n, record_hits=True, fresnel=fresnel)
  b.add_shape(lens, prim.box(2000, 2000, 10), prim.translation(-1000, -1000, 10))
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(3000.0), np.eye(4))
  return b.build()


def cfg(n_rays, seed=5):
  return _abi.CfgArgs(max_ray_length=10000.0, dist_tol=1e-6, max_intersections=50, record_all_hits=True, hit_capacity=60*n_rays, scatter_seed=seed)


def rays_at(theta, n):
  o = np.zeros((n, 3))
  d = np.tile([np.sin(theta), 0.0, np.cos(theta)], (n, 1))
  return o, d


def fresnel_R(theta, n1, n2):
  ci = np.cos(theta); s2 = (n1/n2)**2*(1-ci*ci)
  if s2 >= 1:
    return 1.0
  ct = np.sqrt(1-s2)
  rs, rp = (n1*ci-n2*ct)/(n1*ci+n2*ct), (n1*ct-n2*ci)/(n1*ct+n2*ci)
  return 0.5*(rs*rs+rp*rp)


def first_face_reflected_fraction(hits, n):
  'fraction of rays whose SECOND interaction is not inside the slab: they bounced off its first face'
  second = hits['bounce'] == 1
  return np.count_nonzero(hits['group'][second] == 1)/n


def test_off_is_the_reference_behaviour(oracle):
  n = 2000
  o, d = rays_at(0.4, n)
  r = oracle.trace_rays(slab_scene(False), cfg(n), o, d)
  assert (r['n_segments'] == 3).all() and first_face_reflected_fraction(r['hits'], n) == 0.0


@pytest.mark.parametrize('theta', [0.0, 0.6, np.arctan(N_INDEX), 1.3, 1.5])
def test_reflected_fraction_follows_fresnel(theta, oracle):
  n = 200000
  o, d = rays_at(theta, n)
  r = oracle.trace_rays(slab_scene(True), cfg(n), o, d, threads=0)
  frac = first_face_reflected_fraction(r['hits'], n)
  R = fresnel_R(theta, 1.0, N_INDEX)
  assert abs(frac-R) < 5*np.sqrt(R*(1-R)/n) + 1e-9, (frac, R)
  if theta == 0.0:
    assert abs(R-((N_INDEX-1)/(N_INDEX+1))**2) < 1e-15                       # 4 % for n = 1.5
  if theta == 1.5:
    assert R > 0.6                                                           # towards grazing incidence R -> 1
  # energy conservation: one ray in, one ray out — every ray ends on the absorber, none is lost or doubled
  last = r['hits']['bounce'] == r['n_segments'][r['hits']['ray_index'].astype(np.int64)]-1
  assert np.count_nonzero(last) == n and (r['hits']['group'][last] == 1).all() and (r['final_powers'] == 0).all()
  # multiple internal reflections show up as rays with more than three segments, with the expected first-order weight
  inside_R = fresnel_R(np.arcsin(np.sin(theta)/N_INDEX), N_INDEX, 1.0)
  more = np.count_nonzero(r['n_segments'] > 3)/n
  assert abs(more-(1-R)*inside_R) < 5*np.sqrt(max(more, 1e-6)/n) + 1e-9


def test_total_reflection_is_untouched(oracle):
  'inside the slab beyond the critical angle R = 1: with and without the switch the ray is totally reflected'
  n = 500
  o = np.tile([0.0, 0.0, 15.0], (n, 1))                                      # born inside the slab (medium None -> n1 = 1 at the first face: use a steep exit)
  d = np.tile([np.sin(1.2), 0.0, np.cos(1.2)], (n, 1))
  a = oracle.trace_rays(slab_scene(False), cfg(n), o, d)
  b = oracle.trace_rays(slab_scene(True), cfg(n), o, d)
  assert np.array_equal(a['n_segments'], b['n_segments']) or (b['n_segments'] >= a['n_segments']).all()


@pytest.mark.gpu
@pytest.mark.parametrize('theta', [0.0, 0.9, 1.45])
def test_gpu_matches_oracle_with_fresnel(theta, gpu_engine, oracle):
  n = 50000
  scene = slab_scene(True)
  o, d = rays_at(theta, n)
  c = cfg(n)
  want = oracle.trace_rays(scene, c, o, d, threads=0)
  ds = gpu_engine.scene(scene)
  with ds.trace_rays(c, o, d) as res:
    hits, summary = res.hits(sort=True), res.ray_summary()
  ds.close()
  assert np.array_equal(summary['n_segments'], want['n_segments'])
  for k in ('ray_index', 'bounce', 'group', 'face_id', 'is_entering', 'medium'):
    assert np.array_equal(hits[k], want['hits'][k]), k
  assert np.abs(hits['points']-want['hits']['points']).max() < 1e-8
  assert np.abs(hits['directions']-want['hits']['directions']).max() < 1e-12


@pytest.mark.gpu
def test_gpu_monte_carlo_with_fresnel_matches_oracle(gpu_engine, oracle, sims):
  'a benchmark scene with the switch turned on for its lenses: Monte-Carlo path (seed, source id, ray number key the draws)'
  import copy
  sim = copy.copy(sims('lensesAndMirrors'))
  scene = copy.copy(sim.scene)
  scene.groups = scene.groups.copy()
  scene.groups['fresnel'][scene.groups['optical_type'] == 1] = 1
  sim.scene = scene
  n = 100000
  c = sim.cfg(record_all_hits=True, hit_capacity=12*n)
  want = oracle.trace_mc(scene, sim.source_args(0), c, 0x0DDB1A5E, 1000, n, hit_capacity=12*n, threads=0)
  ds, dsrc = gpu_engine.scene(scene), gpu_engine.source(sim.source_args(0))
  with ds.trace_mc(dsrc, c, 0x0DDB1A5E, 1000, n) as res:
    hits, counts = res.hits(sort=True), res.counts
  ds.close(); dsrc.close()
  assert counts['segments'] == want['counts']['segments'] and counts['hits'] == want['counts']['hits']
  assert counts['segments'] != 7*n                                             # some rays were reflected at a lens face
  for k in ('ray_index', 'bounce', 'group', 'face_id', 'is_entering', 'medium'):
    assert np.array_equal(hits[k], want['hits'][k]), k
  assert np.abs(hits['points']-want['hits']['points']).max() < 1e-8
