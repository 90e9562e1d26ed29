'''
Cases of the traceRay golden (tests/golden/make_traceray_golden.py writes it, tests/test_traceray_golden.py reads it):
benchmark / example scene fixtures with rays drawn from their own light source, and small synthetic scenes that
drive the branches of Ray.traceRay (reference freecad_elements/ray.py:36-281) the shipped scenes do not reach —
total internal reflection, a ray born inside a lens, power decay on lossy mirrors, both grating types, maxIntersections,
and the stochastic surface model (optical_group.py:279-323) on a mirror and on a lens.
'''
import os

import numpy as np

from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.scene_export import primitives as prim
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder
from freecad.optics_design_workbench_b200.simulation.setup import prepare

HERE = os.path.dirname(os.path.abspath(__file__))
SCENES = os.path.join(HERE, 'golden', 'scenes')
SEED = 0x0DDB1A5E

# scene fixture -> number of rays drawn from its first light source (Philox stream SEED, rays 0..n-1)
FIXTURE_CASES = dict(minimal=64, lensesAndMirrors=400, lensesAndMirrorsSequential=400, gettingStarted=200, grating=100)


def _unit(v):
  v = np.asarray(v, dtype=np.float64)
  return v/np.linalg.norm(v, axis=-1, keepdims=True)


def glass_ball():
  'rays born inside an n=1.5 ball (currentMedium is None there, ray.py:85,181-186) and rays grazing it from outside'
  b = SceneBuilder()
  lens = b.add_group('Ball', 'Ball', optical_type='Lens', refractive_index=1.5)
  b.add_shape(lens, prim.sphere(5.0), np.eye(4))
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(40.0), np.eye(4))
  rng = np.random.default_rng(11)
  n = 150
  inside = rng.uniform(-2.5, 2.5, (n, 3))
  d_in = _unit(rng.normal(size=(n, 3)))
  outside = np.column_stack([rng.uniform(-4.9, 4.9, n), rng.uniform(-4.9, 4.9, n), np.full(n, -20.0)])
  d_out = np.tile([0.0, 0.0, 1.0], (n, 1))
  return b.build(), np.vstack([inside, outside]), np.vstack([d_in, d_out]), dict(max_ray_length=200.0, max_intersections=30)


def glass_cube():
  'an n=1.5 cube entered through its bottom face: rays that reach a side wall meet it beyond the critical angle'
  b = SceneBuilder()
  lens = b.add_group('Cube', 'Cube', optical_type='Lens', refractive_index=1.5)
  b.add_shape(lens, prim.box(10, 10, 10), prim.translation(-5, -5, 10))
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(80.0), np.eye(4))
  rng = np.random.default_rng(12)
  n = 200
  o = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.full(n, 5.0)])
  d = _unit(np.column_stack([rng.uniform(-0.8, 0.8, n), rng.uniform(-0.8, 0.8, n), np.ones(n)]))   # whatever reaches a side wall is totally reflected
  return b.build(), o, d, dict(max_ray_length=300.0, max_intersections=25)


def lossy_mirrors():
  'two facing mirrors with Reflectivity 0.5: power halves per bounce until it drops below powerTol (ray.py:280)'
  b = SceneBuilder()
  m1 = b.add_group('MirrorA', 'MirrorA', optical_type='Mirror', reflectivity=0.5, record_hits=True)
  b.add_shape(m1, prim.box(40, 40, 1), prim.translation(-20, -20, 10))
  m2 = b.add_group('MirrorB', 'MirrorB', optical_type='Mirror', reflectivity=0.5)
  b.add_shape(m2, prim.box(40, 40, 1), prim.translation(-20, -20, -11))
  va = b.add_group('Window', 'Window', optical_type='Vacuum', record_hits=True)
  b.add_shape(va, prim.box(40, 40, 0.5), prim.translation(-20, -20, 4))
  rng = np.random.default_rng(13)
  n = 120
  o = np.zeros((n, 3))
  d = _unit(np.column_stack([rng.uniform(-0.08, 0.08, n), rng.uniform(-0.08, 0.08, n), np.ones(n)]))
  return b.build(), o, d, dict(max_ray_length=500.0, max_intersections=100)


def gratings():
  'a reflection grating and a transmission grating slab (ray.py:216-268), three wavelengths'
  b = SceneBuilder()
  gr = b.add_group('ReflGrating', 'ReflGrating', optical_type='Grating', grating_type='Reflection',
                   grating_lines_per_mm=600.0, grating_order=1.0, grating_orientation=(0, 1, 0), record_hits=True)
  b.add_shape(gr, prim.box(30, 30, 2), prim.translation(-15, -15, 20))
  gt = b.add_group('TransGrating', 'TransGrating', optical_type='Grating', grating_type='Transmission', refractive_index=1.4,
                   grating_lines_per_mm=300.0, grating_order=-1.0, grating_orientation=(1, 0, 0), record_hits=True)
  b.add_shape(gt, prim.box(30, 30, 2), prim.translation(-15, -15, -22))
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(90.0), np.eye(4))
  rng = np.random.default_rng(14)
  n = 90
  up = _unit(np.column_stack([rng.uniform(-0.3, 0.3, n), rng.uniform(-0.3, 0.3, n), np.ones(n)]))
  down = up*np.array([1.0, 1.0, -1.0])
  return b.build(), np.zeros((2*n, 3)), np.vstack([up, down]), dict(max_ray_length=300.0, max_intersections=20)


def diffuse_surfaces():
  '''
  applyStochasticRayCorrections: a mirror whose reflected direction is drawn from a density and then modified, and a
  lens whose refracted direction is modified (frosted glass).  The (theta, phi) draws are the engine's Philox draws; the
  rotation formula that turns them into a direction is the reference's.
  '''
  b = SceneBuilder()
  m = b.add_group('Diffuser', 'Diffuser', optical_type='Mirror', reflectivity=0.8, record_hits=True,
                  scatter_density='cos(theta)**2*abs(sin(theta))', power_theta_domain='-pi, -pi/2', power_phi_domain='-pi, pi',
                  modify_density='exp(-theta**2/0.01)*abs(sin(theta))', modify_theta_domain='0, 0.5', modify_phi_domain='0, 2*pi',
                  scatter_resolution=401)
  b.add_shape(m, prim.box(60, 60, 1), prim.translation(-30, -30, 20))
  l = b.add_group('Frosted', 'Frosted', optical_type='Lens', refractive_index=1.5,
                  modify_density='exp(-theta**2/0.02)*abs(sin(theta))', modify_theta_domain='0, 0.6', modify_phi_domain='0, 2*pi',
                  scatter_resolution=401)
  b.add_shape(l, prim.box(60, 60, 4), prim.translation(-30, -30, -24))
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(100.0), np.eye(4))
  rng = np.random.default_rng(15)
  n = 150
  up = _unit(np.column_stack([rng.uniform(-0.4, 0.4, n), rng.uniform(-0.4, 0.4, n), np.ones(n)]))
  down = up*np.array([1.0, 1.0, -1.0])
  return b.build(), np.zeros((2*n, 3)), np.vstack([up, down]), dict(max_ray_length=400.0, max_intersections=40, scatter_seed=77)


SYNTHETIC_CASES = {
  'glass_ball': (glass_ball, (500.0,)),
  'glass_cube': (glass_cube, (500.0,)),
  'lossy_mirrors': (lossy_mirrors, (500.0,)),
  'gratings': (gratings, (450.0, 633.0, 1000.0)),
  'diffuse_surfaces': (diffuse_surfaces, (500.0,)),
}


def fixture_case(name):
  'PreparedSimulation of a scene fixture'
  return prepare(os.path.join(SCENES, name+'.npz'))


def synthetic_cfg(settings, **overrides):
  return _abi.CfgArgs(dist_tol=0.01, power_tol=1e-6, sequential=False, **{**settings, **overrides})
