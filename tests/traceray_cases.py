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


# ---- cases for the decisive branches of Ray.findNearestIntersection (ray.py:290-452) and find.relevantOpticalObjects ----
# The settings dict of these cases may carry extras that are not trace settings (popped by split_settings):
#   sequence  [[group, ...], ...]  SequentialModeElements_NN lists (SequentialMode on)
#   ignored   [group, ...]          IgnoredOpticalElements of the light source
#   frames    {group: (T, S)}       the reference side sees the group's Shape in its own coordinates (own placement S) and
#                                   reaches the world through gpM = T (ray.py:338-339: M = gpM * pMi)
TOL = 0.01          # DistanceTolerance of the synthetic cases (synthetic_cfg)


def slab_stack():
  '''
  Lens slabs along +z separated by gaps of 0, 0.5, 1.5, 2.5, 3, 5.5 and 6.5 tolerances: coincident faces (prefer the group
  that is not the current medium), a second hit inside minDist + 2 tol, hits that are found but dropped by that filter,
  hits inside / outside the maxRayLength + 5 tol window of the previous one.  An enclosing absorber sphere is the FIRST
  shell candidate (bbox distance 0), so every nearer hit is found after maxRayLength has already been shrunk.
  '''
  b = SceneBuilder()
  z, gaps = 10.0, [0.0, 0.5, 1.5, 2.5, 3.0, 5.5, 6.5]
  for i in range(len(gaps)+1):
    g = b.add_group(f'Slab{i}', f'Slab{i}', optical_type='Lens', refractive_index=1.3+0.05*i, record_hits=True)
    b.add_shape(g, prim.box(30, 30, 4), prim.translation(-15, -15, z))
    z += 4.0 + (gaps[i]*TOL if i < len(gaps) else 0.0)
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(90.0), np.eye(4))
  rng = np.random.default_rng(21)
  n = 160
  o = np.column_stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), np.zeros(n)])
  d = _unit(np.column_stack([rng.uniform(-0.15, 0.15, n), rng.uniform(-0.15, 0.15, n), np.ones(n)]))
  d[:20] = [0.0, 0.0, 1.0]                                 # exactly perpendicular: the gaps are met at exactly their width
  return b.build(), o, d, dict(max_ray_length=300.0, max_intersections=40)


def behind_the_start():
  '''
  Quirk Q4: BoundBox.intersect and Curve.intersect work on the infinite LINE, so objects behind the start pass the culls
  and produce crossing points; only "distance to the forward segment < tol" removes them.  A mirror box, a lens ball and
  a torus sit on the backward half of the line; rays also start inside a lens box (crossings of that shell on both sides).
  '''
  b = SceneBuilder()
  m = b.add_group('BackMirror', 'BackMirror', optical_type='Mirror', record_hits=True)
  b.add_shape(m, prim.box(20, 20, 2), prim.translation(-10, -10, -12))
  l = b.add_group('BackBall', 'BackBall', optical_type='Lens', refractive_index=1.5, record_hits=True)
  b.add_shape(l, prim.sphere(3.0), prim.translation(0, 0, -20))
  t = b.add_group('BackRing', 'BackRing', optical_type='Absorber', record_hits=True)
  b.add_shape(t, prim.torus(6.0, 2.0), prim.translation(0, 0, -30))
  box = b.add_group('Block', 'Block', optical_type='Lens', refractive_index=1.5, record_hits=True)
  b.add_shape(box, prim.box(10, 10, 10), prim.translation(-5, -5, 20))
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(80.0), np.eye(4))
  rng = np.random.default_rng(22)
  n = 100
  o1 = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.zeros(n)])
  d1 = _unit(np.column_stack([rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n), np.ones(n)]))
  o2 = np.column_stack([rng.uniform(-4, 4, n), rng.uniform(-4, 4, n), rng.uniform(21, 29, n)])      # born inside the block
  d2 = _unit(rng.normal(size=(n, 3)))
  return b.build(), np.vstack([o1, o2]), np.vstack([d1, d2]), dict(max_ray_length=200.0, max_intersections=30)


def end_of_the_ray():
  '''
  Hits around the end of the ray: rule (ii) of ray.py:424-426 accepts a crossing up to one tolerance BEYOND
  start + maxRayLength (distance to the finite segment), provided the enlarged bounding box is nearer than maxRayLength.
  A ball whose surface is met between maxRayLength - 3 tol and maxRayLength + 3 tol (its box starts well before), and a
  plate (box distance = hit distance - tol).
  '''
  b = SceneBuilder()
  ball = b.add_group('Ball', 'Ball', optical_type='Absorber', record_hits=True)
  b.add_shape(ball, prim.sphere(5.0), prim.translation(0, 0, 55.0))
  plate = b.add_group('Plate', 'Plate', optical_type='Absorber', record_hits=True)
  b.add_shape(plate, prim.box(10, 10, 1), prim.translation(15, -5, 50.0))
  rng = np.random.default_rng(23)
  n = 120
  # ball: off-axis by 3 -> surface at z = 55 - 4 = 51; start so that the crossing lies at 50 + k tol, k in (-3, 3)
  k = rng.uniform(-3, 3, n)
  o1 = np.column_stack([np.full(n, 3.0), np.zeros(n), 51.0 - 50.0 - k*TOL])
  o2 = np.column_stack([np.full(n, 20.0), rng.uniform(-4, 4, n), -k*TOL])                            # plate at distance 50 + k tol
  d = np.tile([0.0, 0.0, 1.0], (2*n, 1))
  return b.build(), np.vstack([o1, o2]), d, dict(max_ray_length=50.0, max_intersections=5)


def sequence_and_ignore_list():
  '''
  find.relevantOpticalObjects (find.py:79-104): IgnoredOpticalElements of the light source and the sequential filter —
  a step with two groups, a group that appears in two steps, and the end of the list (nothing is hittable any more).
  '''
  b = SceneBuilder()
  m1 = b.add_group('M1', 'M1', optical_type='Mirror', record_hits=True)
  b.add_shape(m1, prim.box(40, 40, 1), prim.translation(-20, -20, 30))
  m2 = b.add_group('M2', 'M2', optical_type='Mirror', record_hits=True)
  b.add_shape(m2, prim.box(40, 40, 1), prim.translation(-20, -20, -31))
  win = b.add_group('Window', 'Window', optical_type='Vacuum', record_hits=True)
  b.add_shape(win, prim.box(40, 40, 2), prim.translation(-20, -20, 10))
  blocker = b.add_group('Blocker', 'Blocker', optical_type='Absorber', record_hits=True)        # on the ignore list
  b.add_shape(blocker, prim.box(40, 40, 1), prim.translation(-20, -20, 5))
  lens = b.add_group('Ball', 'Ball', optical_type='Lens', refractive_index=1.5, record_hits=True)
  b.add_shape(lens, prim.sphere(4.0), prim.translation(0, 0, -15))
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(100.0), np.eye(4))
  rng = np.random.default_rng(24)
  n = 150
  o = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.zeros(n)])
  d = _unit(np.column_stack([rng.uniform(-0.05, 0.05, n), rng.uniform(-0.05, 0.05, n), np.ones(n)]))
  # window (entry + exit: two steps), mirror 1, window again on the way back, ball or mirror 2, ball, mirror 2 — then nothing
  sequence = [[win], [win], [m1], [win], [win], [lens, m2], [lens, m2], [m2, m1]]
  return b.build(sequence=sequence), o, d, dict(max_ray_length=300.0, max_intersections=40, sequence=sequence, ignored=[blocker])


def placed_groups():
  '''
  Groups reached through placements (ray.py:332-352): the reference transforms the RAY into the coordinates of the group's
  Shape (gpMi = pM * gpMi) and the hit point and normal back (gpM = gpM * pMi); the engine's scene export folds the same
  matrices into the faces.  Every group here has its own placement S inside a rotated + shifted container T.
  '''
  T1 = prim.translation(3, -2, 40) @ prim.rotation((1, 1, 0), 0.6)
  T2 = prim.translation(-4, 5, -35) @ prim.rotation((0, 1, 1), -0.8)
  S1 = prim.translation(1, 2, 3) @ prim.rotation((0, 0, 1), 0.3)
  S2 = prim.translation(-2, 0, 1) @ prim.rotation((1, 0, 0), 1.1)
  parts = [('Prism', dict(optical_type='Lens', refractive_index=1.6, record_hits=True), prim.box(12, 12, 6), T1, S1, prim.translation(-6, -6, -3)),
           ('Dish', dict(optical_type='Mirror', reflectivity=0.9, record_hits=True), prim.plano_convex_lens(20.0, 8.0, 1.0), T2, S2, np.eye(4)),
           ('Shell', dict(optical_type='Absorber', record_hits=True), prim.sphere(120.0), np.eye(4), np.eye(4), np.eye(4))]
  world, local, frames = SceneBuilder(), SceneBuilder(), {}
  for name, props, faces, T, S, own in parts:
    gw = world.add_group(name, name, **props)
    world.add_shape(gw, faces, T @ own)                    # world = T * local (the group's own placement S cancels: gpM * pMi * S)
    gl = local.add_group(name, name, **props)
    local.add_shape(gl, faces, S @ own)                    # what group.Shape holds: own placement applied
    frames[gw] = (T, S)
  rng = np.random.default_rng(25)
  n = 150
  up = _unit(np.column_stack([rng.uniform(-0.12, 0.12, n), rng.uniform(-0.12, 0.12, n), np.ones(n)]))
  down = up*np.array([1.0, 1.0, -1.0])
  return world.build(), np.zeros((2*n, 3)), np.vstack([up, down]), dict(max_ray_length=400.0, max_intersections=30, frames=frames, shape_scene=local.build())


def curved_surfaces():
  'every elementary surface kind with its trims: cylinder, cone frustum, torus, sphere, lens cap, discs'
  b = SceneBuilder()
  cyl = b.add_group('Rod', 'Rod', optical_type='Lens', refractive_index=1.45, record_hits=True)
  b.add_shape(cyl, prim.cylinder(3.0, 12.0), prim.translation(0, 0, 14) @ prim.rotation((1, 0, 0), 0.9))
  cone = b.add_group('Frustum', 'Frustum', optical_type='Mirror', reflectivity=0.95, record_hits=True)
  b.add_shape(cone, prim.cone(5.0, 2.0, 8.0), prim.translation(9, 3, -20) @ prim.rotation((0, 1, 0), 0.4))
  ring = b.add_group('Ring', 'Ring', optical_type='Lens', refractive_index=1.7, record_hits=True)
  b.add_shape(ring, prim.torus(8.0, 2.5), prim.translation(-6, 0, 32) @ prim.rotation((1, 1, 0), 0.5))
  lens = b.add_group('Singlet', 'Singlet', optical_type='Lens', refractive_index=1.5, record_hits=True)
  b.add_shape(lens, prim.plano_convex_lens(15.0, 6.0, 1.0), prim.translation(0, 0, -40))
  ab = b.add_group('Shell', 'Shell', optical_type='Absorber', record_hits=True)
  b.add_shape(ab, prim.sphere(110.0), np.eye(4))
  rng = np.random.default_rng(26)
  n = 150
  centres = np.array([[0, -4.7, 17.7], [10.6, 3, -16.3], [-6, 0, 32], [0, 0, -39]], dtype=np.float64)     # rod, frustum, ring, singlet
  spread = np.array([5.0, 5.0, 10.0, 6.0])
  o = rng.uniform(-1.0, 1.0, (4*n, 3))
  aim = np.repeat(centres, n, axis=0) + rng.uniform(-1, 1, (4*n, 3))*np.repeat(spread, n)[:, None]
  d = _unit(aim-o)
  return b.build(), o, d, dict(max_ray_length=400.0, max_intersections=40)


SYNTHETIC_CASES = {
  'glass_ball': (glass_ball, (500.0,)),
  'glass_cube': (glass_cube, (500.0,)),
  'lossy_mirrors': (lossy_mirrors, (500.0,)),
  'gratings': (gratings, (450.0, 633.0, 1000.0)),
  'diffuse_surfaces': (diffuse_surfaces, (500.0,)),
  'slab_stack': (slab_stack, (500.0,)),
  'behind_the_start': (behind_the_start, (500.0,)),
  'end_of_the_ray': (end_of_the_ray, (500.0,)),
  'sequence_and_ignore_list': (sequence_and_ignore_list, (500.0,)),
  'placed_groups': (placed_groups, (500.0,)),
  'curved_surfaces': (curved_surfaces, (500.0,)),
}


def fixture_case(name):
  'PreparedSimulation of a scene fixture'
  return prepare(os.path.join(SCENES, name+'.npz'))


def split_settings(settings):
  'trace settings, and the extras of a synthetic case that are not trace settings'
  settings = dict(settings)
  extras = {k: settings.pop(k, None) for k in ('sequence', 'ignored', 'frames', 'shape_scene')}
  return settings, extras


def synthetic_cfg(settings, **overrides):
  settings, extras = split_settings(settings)
  return _abi.CfgArgs(dist_tol=TOL, power_tol=1e-6, sequential=bool(extras['sequence']), **{**settings, **overrides})
