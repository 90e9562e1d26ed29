'''
Stochastic surface model (reference freecad_elements/optical_group.py:212-323 applyStochasticRayCorrections):
tabulated (theta, phi) densities per optical group, the rotation formula, "no change" DiracDelta densities, refused
per-hit densities; on the GPU the same draws as the oracle.  Scene: test/50-old-tests/mirror-diffuse.FCStd.
'''
import numpy as np
import pytest
from scipy import stats

from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.distributions import scatter_tables
from freecad.optics_design_workbench_b200.scene_export import primitives as prim
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder

from conftest import SEED


def diffuse_mirror_scene(density='cos(theta)**2*abs(sin(theta))', theta_dom='-pi, -pi/2', phi_dom='-pi, pi', modify=''):
  b = SceneBuilder()
  m = b.add_group('Mirror', 'Mirror', optical_type='Mirror', reflectivity=0.5, record_hits=True,
                  scatter_density=density, power_theta_domain=theta_dom, power_phi_domain=phi_dom,
                  modify_density=modify, scatter_resolution=401)
  b.add_shape(m, prim.box(200, 200, 1), prim.translation(-100, -100, 10))
  a = b.add_group('Abs', 'Abs', optical_type='Absorber', record_hits=True)
  b.add_shape(a, prim.sphere(60.0), np.eye(4))           # the source sits inside: everything ends on the sphere
  return b.build()


def rays(n, tilt=0.3):
  d = np.tile([np.sin(tilt), 0.0, np.cos(tilt)], (n, 1))
  return np.zeros((n, 3)), d


def test_density_classification():
  assert scatter_tables('', '-pi/2, pi/2', '0, 2*pi') is None
  assert scatter_tables('DiracDelta(theta)', '-pi/2, pi/2', '0, 2*pi') is None               # lens-optimizer / playground scenes
  assert scatter_tables('DiracDelta(theta)*DiracDelta(phi)', '-pi/2, pi/2', '0, 2*pi') is None   # mirror / mirror-diffuse absorbers
  t = scatter_tables('cos(theta)**2 * abs(sin(theta))', '-pi, -pi/2', '-pi,pi')
  assert t.n_rows == 1 and t.first_cdf.shape[1] == 1005 and t.first_domain == (-np.pi, -np.pi/2)
  with pytest.raises(NotImplementedError):
    scatter_tables('exp(-(theta-theta_refl)**2)', '-pi/2, pi/2', '0, 2*pi')
  with pytest.raises(NotImplementedError):
    scatter_tables('DiracDelta(theta-0.1)', '-pi/2, pi/2', '0, 2*pi')


def test_ideal_surfaces_are_unchanged_by_no_op_densities(oracle):
  o_, d_ = rays(200)
  a = diffuse_mirror_scene(density='', modify='')
  b = diffuse_mirror_scene(density='', modify='DiracDelta(theta)')
  assert not a.scatters and not b.scatters
  cfg = _abi.CfgArgs(max_ray_length=500, record_all_hits=True)
  ra, rb = oracle.trace_rays(a, cfg, o_, d_), oracle.trace_rays(b, cfg, o_, d_)
  np.testing.assert_array_equal(ra['hits']['points'], rb['hits']['points'])
  # specular: second hit direction is the mirror image
  h = ra['hits']
  np.testing.assert_allclose(h['directions'][h['bounce'] == 1], [[np.sin(0.3), 0, -np.cos(0.3)]]*200, atol=1e-12)


def test_diffuse_mirror_angles_follow_the_density(oracle):
  'theta measured from the normal that points along the propagation (into the mirror): domain (-pi, -pi/2) sends rays back'
  n = 200000
  o_, d_ = rays(n)
  sc = diffuse_mirror_scene()
  cfg = _abi.CfgArgs(max_ray_length=500, record_all_hits=True, scatter_seed=1234, hit_capacity=3*n)
  r = oracle.trace_rays(sc, cfg, o_, d_, hit_capacity=3*n, threads=0)
  h = r['hits']
  out = h['directions'][h['bounce'] == 1]                 # direction after the mirror, recorded at the absorber hit
  assert len(out) == n
  np.testing.assert_allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-12)
  nrm = np.array([0.0, 0.0, 1.0])                         # mirror face normal flipped along the propagation
  cos_t = out@nrm
  assert cos_t.max() < 0                                  # all rays leave the mirror
  theta = -np.arccos(cos_t)                               # in (-pi, -pi/2): cos(theta) = cos_t
  # density cos^2 |sin| on (-pi, -pi/2):  CDF(theta) = 1 + cos^3(theta)  ->  u = 1 + cos_t^3 is uniform
  assert stats.kstest(1+cos_t**3, stats.uniform(0, 1).cdf).pvalue > 1e-3
  # azimuth about the normal, measured from a x n with a = n x d_in: uniform over (-pi, pi)
  a = np.cross(nrm, d_[0]); a /= np.linalg.norm(a)
  axn = np.cross(a, nrm)
  phi = np.arctan2(out@a, out@axn)
  # sin(theta) < 0 flips the azimuth by pi; uniformity is what is pinned here
  assert stats.kstest(phi, stats.uniform(-np.pi, 2*np.pi).cdf).pvalue > 1e-3
  powers = h['powers'][h['bounce'] == 1]
  np.testing.assert_allclose(powers, 0.5)                 # Reflectivity still applies


def test_scatter_draws_depend_on_seed_ray_and_bounce_only(oracle):
  o_, d_ = rays(500)
  sc = diffuse_mirror_scene()
  cfg1 = _abi.CfgArgs(max_ray_length=500, record_all_hits=True, scatter_seed=7)
  cfg2 = _abi.CfgArgs(max_ray_length=500, record_all_hits=True, scatter_seed=8)
  a, b, c = oracle.trace_rays(sc, cfg1, o_, d_), oracle.trace_rays(sc, cfg1, o_, d_, threads=0), oracle.trace_rays(sc, cfg2, o_, d_)
  np.testing.assert_array_equal(a['hits']['directions'], b['hits']['directions'])
  assert np.abs(a['hits']['directions']-c['hits']['directions']).max() > 0.1


def test_reference_mirror_diffuse_scene_imports_and_runs(oracle, sims):
  sim = sims('mirrorDiffuse')
  assert len(sim.scene.scatters) == 1 and sim.scene.group_scatter.tolist()[0] == [0, -1]
  r = oracle.trace_mc(sim.scene, sim.source_args(0), sim.cfg(), SEED, 0, 5000, threads=0)
  assert r['counts']['hits'] > 0 and r['counts']['depth_terminated'] == 0


# ------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_gpu_diffuse_mirror_equals_oracle(gpu_engine, oracle, sims):
  n = 50000
  o_, d_ = rays(n)
  for sc in (diffuse_mirror_scene(), diffuse_mirror_scene(modify='exp(-theta**2/0.01)*abs(sin(theta))')):
    cfg = _abi.CfgArgs(max_ray_length=500, record_all_hits=True, scatter_seed=99, hit_capacity=4*n)
    with gpu_engine.scene(sc).trace_rays(cfg, o_, d_) as res:
      gc, gh = res.counts, res.hits(sort=True)
    o = oracle.trace_rays(sc, cfg, o_, d_, hit_capacity=4*n, threads=0)
    assert gc == o['counts']
    np.testing.assert_array_equal(gh['face_id'], o['hits']['face_id'])
    np.testing.assert_allclose(gh['points'], o['hits']['points'], rtol=0, atol=1e-8)
    np.testing.assert_allclose(gh['directions'], o['hits']['directions'], rtol=0, atol=1e-9)
  sim = sims('mirrorDiffuse')
  sa = sim.source_args(0)
  cfg = sim.cfg(record_all_hits=True, hit_capacity=10*n)
  with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sa), cfg, SEED, 0, n) as res:
    gc, gh = res.counts, res.hits(sort=True)
  o = oracle.trace_mc(sim.scene, sa, cfg, SEED, 0, n, hit_capacity=10*n, threads=0)
  assert gc == o['counts']
  np.testing.assert_allclose(gh['points'], o['hits']['points'], rtol=0, atol=1e-8)
