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
    scatter_tables('DiracDelta(theta-0.1)', '-pi/2, pi/2', '0, 2*pi')
  # densities with per-hit parameters become a family over the incidence angle (one family for a mirror, two for a lens) ...
  f = scatter_tables('exp(-(theta-theta_refl)**2/0.02)', 'pi/2, pi', '0, 2*pi', resolution=201, param_tables=11)
  assert f.n_tables == 11 and f.first_cdf.shape == (11, 1, 201) and f.phi_cdf.shape == (11, 3)
  g = scatter_tables('exp(-(theta-theta_refl)**2/0.02)', '0, pi/2', '0, 2*pi', resolution=201, optical_type='Lens', refractive_index=1.5, param_tables=11)
  assert g.n_tables == 22
  with pytest.raises(NotImplementedError):            # ... unless they also depend on phi
    scatter_tables('exp(-(theta-theta_refl)**2)*(1+cos(phi))', '-pi/2, pi/2', '0, 2*pi')


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


# ---- densities the reference re-compiles for every hit (optical_group.py:288-307) -------------------------------------
LOBE = 'exp(-(theta-theta_refl)**2/0.02)'            # a Gaussian lobe around the specular direction, whatever the incidence


def lobe_mirror_scene():
  b = SceneBuilder()
  m = b.add_group('Mirror', 'Mirror', optical_type='Mirror', record_hits=True, scatter_density=LOBE,
                  power_theta_domain='pi/2, pi', power_phi_domain='0, 2*pi', scatter_resolution=801)
  b.add_shape(m, prim.box(400, 400, 1), prim.translation(-200, -200, 10))
  a = b.add_group('Abs', 'Abs', optical_type='Absorber', record_hits=True)
  b.add_shape(a, prim.sphere(90.0), np.eye(4))
  return b.build()


def outgoing_theta(hits):
  'angle between the direction AFTER the mirror (incoming direction of the second hit) and the normal along the propagation (+z)'
  d = hits['directions'][hits['bounce'] == 1]
  return np.arccos(np.clip(d[:, 2]/np.linalg.norm(d, axis=1), -1, 1))


@pytest.mark.parametrize('tilt', [0.05, 0.35, 0.9, 1.2])     # (at exactly normal incidence the rotation axis n x d vanishes and nothing is rotated, as in the reference)
def test_per_hit_density_follows_the_incidence_angle(tilt, oracle):
  '''
  the lobe sits at theta_refl = pi - theta_in for every incidence: the outgoing angle (measured from the normal that
  points into the mirror) is distributed like the density evaluated at the hit's own theta_in — here the analytic
  truncated Gaussian around pi - tilt on [pi/2, pi] — up to the half-degree grid of the family
  '''
  n = 40000
  scene = lobe_mirror_scene()
  assert scene.scatters[0].n_tables == 91
  o_, d_ = rays(n, tilt=tilt)
  r = oracle.trace_rays(scene, _abi.CfgArgs(max_ray_length=500, record_all_hits=True, scatter_seed=5), o_, d_, threads=0)
  th = outgoing_theta(r['hits'])
  assert len(th) == n
  grid = np.linspace(np.pi/2, np.pi, 20001)
  k = int(round(tilt/(np.pi/2)*90))                              # the family member the hit uses
  centre = np.pi - k*(np.pi/2)/90
  pdf = np.exp(-(grid-centre)**2/0.02)
  cdf = np.concatenate([[0], np.cumsum((pdf[1:]+pdf[:-1])/2)]); cdf /= cdf[-1]
  assert stats.kstest(th, lambda x: np.interp(x, grid, cdf)).pvalue > 0.01
  assert abs(centre-(np.pi-tilt)) <= 0.5*np.pi/180+1e-12        # never more than half a degree off the hit's own angle


@pytest.mark.reference
def test_per_hit_density_against_the_reference_sampler(oracle):
  'the reference compiles the density with the hit\'s parameters and draws (optical_group.py:307-308): same distribution of theta'
  import reference_shim
  ref = reference_shim.load()
  tilt, n = 0.35, 40000
  vrv = ref.distributions.VectorRandomVariable(probabilityDensity='('+LOBE+')', variableOrder=('theta', 'phi'),
                                               variableDomains=dict(theta=(np.pi/2, np.pi), phi=(0, 2*np.pi)))
  vrv.compile(theta_in=tilt, phi_in=0, theta_refl=np.pi-tilt, phi_refl=0)
  np.random.seed(7)
  want, _ = vrv.draw(N=n)
  o_, d_ = rays(n, tilt=tilt)
  r = oracle.trace_rays(lobe_mirror_scene(), _abi.CfgArgs(max_ray_length=500, record_all_hits=True, scatter_seed=6), o_, d_, threads=0)
  got = outgoing_theta(r['hits'])
  # two-sample test; the family member for 0.35 rad is 0.349 rad (0.05 degrees off)
  assert stats.ks_2samp(got, np.asarray(want, dtype=float)).pvalue > 0.01


# ------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_gpu_per_hit_density_equals_oracle(gpu_engine, oracle):
  n = 30000
  scene = lobe_mirror_scene()
  rng = np.random.default_rng(3)
  o_ = np.zeros((n, 3))
  d_ = np.column_stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1.5, 1.5, n), np.ones(n)])         # incidence angles from 0 to 65 degrees
  cfg = _abi.CfgArgs(max_ray_length=500, record_all_hits=True, scatter_seed=11, hit_capacity=4*n)
  want = oracle.trace_rays(scene, cfg, o_, d_, hit_capacity=4*n, threads=0)
  ds = gpu_engine.scene(scene)
  with ds.trace_rays(cfg, o_, d_) as res:
    gh = res.hits(sort=True)
  ds.close()
  np.testing.assert_array_equal(gh['face_id'], want['hits']['face_id'])
  np.testing.assert_allclose(gh['points'], want['hits']['points'], rtol=0, atol=1e-7)
  np.testing.assert_allclose(gh['directions'], want['hits']['directions'], rtol=0, atol=1e-9)

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
