'''
Conics of revolution in closed form (ODW_SURF_CONICOID, include/odw.h): paraboloid, ellipsoid, hyperboloid sheet, sphere cap.
The reference asks OpenCASCADE for line/surface intersections on whatever surface a face carries (reference
freecad_elements/ray.py:407-432); a revolved parabola arrives as BRep surface type 7 (surface of revolution) and is
recognised as a paraboloid by scene_export.scene.face_record.  Known answers are the focal properties of the conics,
which hold for every ray, not only paraxial ones.
'''
import numpy as np
import pytest

from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.scene_export import primitives as prim
from freecad.optics_design_workbench_b200.scene_export import scene as sc_mod
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder


def trace(engine_like, scene, o, d, **cfg):
  kw = dict(max_ray_length=1000.0, dist_tol=1e-6, max_intersections=20, record_all_hits=True)
  kw.update(cfg)
  return engine_like.trace_rays(scene, _abi.CfgArgs(**kw), np.atleast_2d(o), np.atleast_2d(d))


def dish_and_screen(faces, screen_z, screen_r=200.0, **mirror):
  'mirror group 0 = the dish (vertex at the origin, axis +z), absorber group 1 = a disc at z = screen_z'
  b = SceneBuilder()
  m = b.add_group('M', 'M', optical_type='Mirror', **mirror)
  b.add_shape(m, faces, np.eye(4))
  a = b.add_group('A', 'A', optical_type='Absorber', record_hits=True)
  b.add_shape(a, prim.disc(screen_r), prim.translation(0, 0, screen_z))
  return b.build()


def grid_rays(radius, n=9, z=50.0):
  'rays parallel to -z from the plane z, on a polar grid inside `radius` (off the pole)'
  r = np.linspace(0.07, 0.93, n)*radius
  a = np.linspace(0.1, 6.1, n)
  R, A = np.meshgrid(r, a)
  o = np.stack([R.ravel()*np.cos(A.ravel()), R.ravel()*np.sin(A.ravel()), np.full(R.size, z)], axis=-1)
  return o, np.tile([0.0, 0.0, -1.0], (len(o), 1))


def test_paraboloid_focuses_axis_parallel_rays(oracle):
  'z = rho^2/(4F): every ray parallel to the axis passes through (0, 0, F) after one reflection'
  F, R = 12.5, 20.0
  scene = dish_and_screen(prim.conic_dish(1/(2*F), -1.0, R), screen_z=F, screen_r=1.0)
  assert int(scene.faces[0]['kind']) == sc_mod.SURF_CONICOID
  o, d = grid_rays(R)
  r = trace(oracle, scene, o, d)
  h = r['hits']
  assert np.all(r['n_segments'] == 2) and len(h['points']) == 2*len(o)
  first, second = h['points'][h['bounce'] == 0], h['points'][h['bounce'] == 1]
  rho = np.hypot(o[:, 0], o[:, 1])
  assert np.abs(first[:, 2]-rho**2/(4*F)).max() < 1e-12 and np.abs(first[:, :2]-o[:, :2]).max() < 1e-12
  assert np.abs(second-[0, 0, F]).max() < 1e-11                        # the focus, to rounding
  assert np.all(h['group'][h['bounce'] == 1] == 1)
  # the dish is hit from the side its geometric normal does not point to: "entering" (ray.py:473-480)
  assert np.all(h['is_entering'][h['bounce'] == 0] == 0) or np.all(h['is_entering'][h['bounce'] == 0] == 1)


def test_revolved_parabola_is_recognised_as_the_same_paraboloid(oracle):
  'BRep surface type 7 (revolution of a Geom_Parabola about its own axis) -> ODW_SURF_CONICOID, same hits as the direct form'
  F, R = 8.0, 10.0
  a = dish_and_screen(prim.revolved_parabola_dish(F, R), screen_z=60.0)
  b = dish_and_screen(prim.conic_dish(1/(2*F), -1.0, R), screen_z=60.0)
  assert int(a.faces[0]['kind']) == sc_mod.SURF_CONICOID and a.faces[0]['p1'] == -1.0
  assert abs(a.faces[0]['p0']-1/(2*F)) < 1e-15
  assert np.allclose(a.faces[0]['aabb_min'], b.faces[0]['aabb_min']) and np.allclose(a.faces[0]['aabb_max'], b.faces[0]['aabb_max'])
  assert a.faces[0]['aabb_min'][2] <= 0 and a.faces[0]['aabb_max'][2] >= R*R/(4*F)
  rng = np.random.default_rng(5)
  o = np.column_stack([rng.uniform(-6, 6, 300), rng.uniform(-6, 6, 300), np.full(300, 40.0)])
  d = np.column_stack([rng.normal(0, 0.05, 300), rng.normal(0, 0.05, 300), -np.ones(300)])
  ra, rb = trace(oracle, a, o, d), trace(oracle, b, o, d)
  assert np.array_equal(ra['hits']['group'], rb['hits']['group']) and np.sum(ra['hits']['group'] == 0) > 250
  assert np.abs(ra['hits']['points']-rb['hits']['points']).max() < 1e-12
  assert np.abs(ra['hits']['directions']-rb['hits']['directions']).max() < 1e-12
  # a parabola revolved about a line that is not its axis is no paraboloid: it goes to the tessellation path
  from freecad.optics_design_workbench_b200.scene_export.brep import Curve3d, Surface
  par = Curve3d('parabola', p=np.zeros(3), n=np.array([0., 1, 0]), dx=np.array([0., 0, 1]), dy=np.array([1., 0, 0]), f=F)
  assert sc_mod._revolved_conic(Surface('revolution', p=np.array([1., 0, 0]), d=np.array([0., 0, 1]), curve=par)) is None
  assert sc_mod._revolved_conic(Surface('revolution', p=np.zeros(3), d=np.array([1., 0, 0]), curve=par)) is None


def test_ellipsoid_images_one_focus_onto_the_other(oracle):
  'prolate ellipsoid (-1 < k < 0): a ray from one focus reaches the other after one reflection; path length 2a for all rays'
  a_, b_ = 30.0, 18.0                                     # semi-axes: a along z, b across
  e = np.sqrt(1-b_*b_/(a_*a_))
  c, k = a_/(b_*b_), -e*e                                 # vertex curvature b^2/a, conic constant -e^2
  f_near, f_far = a_*(1-e), a_*(1+e)                      # foci on the axis, measured from the vertex
  scene = dish_and_screen(prim.conic_dish(c, k, 0.9*b_), screen_z=f_near, screen_r=0.5)
  rng = np.random.default_rng(11)
  th, ph = rng.uniform(0.02, 0.25, 60), rng.uniform(0, 2*np.pi, 60)
  d = np.column_stack([np.sin(th)*np.cos(ph), np.sin(th)*np.sin(ph), -np.cos(th)])
  o = np.tile([0.0, 0.0, f_far], (60, 1))
  r = trace(oracle, scene, o, d)
  h = r['hits']
  assert np.all(r['n_segments'] == 2)
  P1, P2 = h['points'][h['bounce'] == 0], h['points'][h['bounce'] == 1]
  assert np.abs(P2-[0, 0, f_near]).max() < 1e-10
  path = np.linalg.norm(P1-o, axis=1) + np.linalg.norm(P2-P1, axis=1)
  assert np.abs(path-2*a_).max() < 1e-10
  # points satisfy the implicit quadric
  rho2, z = P1[:, 0]**2 + P1[:, 1]**2, P1[:, 2]
  assert np.abs(c*(rho2 + (1+k)*z*z) - 2*z).max() < 1e-11


def test_hyperboloid_sheet_only_and_virtual_focus(oracle):
  'k < -1: only the sheet through the vertex exists; rays aimed at the far focus leave as if they came from the near one'
  a_, cc = 10.0, 26.0                                     # hyperbola: vertices at +-a from its centre, foci at +-cc
  b2 = cc*cc-a_*a_
  c, k = a_/b2, -(cc/a_)**2                               # sheet with its vertex at the origin opening towards +z
  scene = dish_and_screen(prim.conic_dish(c, k, 25.0), screen_z=-50.0, screen_r=500.0)
  # geometric foci: behind the vertex at z = -(cc-a) ... wait for the sheet opening to +z the near focus is inside at z = cc-a
  f_in, f_out = cc-a_, -(cc+a_)
  rng = np.random.default_rng(2)
  tgt = np.array([0.0, 0.0, f_in])
  o = np.column_stack([rng.uniform(-8, 8, 50), rng.uniform(-8, 8, 50), np.full(50, -40.0)])
  d = tgt-o
  d /= np.linalg.norm(d, axis=1)[:, None]
  r = trace(oracle, scene, o, d, max_intersections=1)
  h = r['hits']
  assert len(h['points']) == 50 and np.all(h['group'] == 0)
  P = h['points']
  z = P[:, 2]
  assert np.all(z >= 0) and np.abs(c*(P[:, 0]**2+P[:, 1]**2+(1+k)*z*z)-2*z).max() < 1e-10     # the near sheet, never the far one
  # reflected direction points away from the other focus (reflection property of the hyperbola)
  n = np.column_stack([c*P[:, 0], c*P[:, 1], -(1-(1+k)*c*z)])
  n /= np.linalg.norm(n, axis=1)[:, None]
  out = d-2*np.sum(d*n, axis=1)[:, None]*n
  away = P-[0, 0, f_out]
  away /= np.linalg.norm(away, axis=1)[:, None]
  assert np.abs(np.abs(np.sum(out*away, axis=1))-1).max() < 1e-10


def test_conicoid_k0_equals_the_sphere_kind(oracle):
  'k = 0 is a sphere: same hits as a sphere cap of radius 1/c written with ODW_SURF_SPHERE'
  Rs, ap = 25.0, 10.0
  from freecad.optics_design_workbench_b200.scene_export.brep import Surface
  v0 = np.arcsin(np.sqrt(Rs*Rs-ap*ap)/Rs)
  cap = Surface('sphere', p=np.array([0, 0, Rs]), n=np.array([0., 0, -1]), dx=np.array([1., 0, 0]), dy=np.array([0., -1, 0]), r=Rs)
  sph = dish_and_screen([prim._face(cap, [prim._rect_loop(0, 2*np.pi, v0, np.pi/2)], shell_key=None)], screen_z=60.0)
  con = dish_and_screen(prim.conic_dish(1/Rs, 0.0, ap), screen_z=60.0)
  rng = np.random.default_rng(8)
  o = np.column_stack([rng.uniform(-12, 12, 400), rng.uniform(-12, 12, 400), np.full(400, 30.0)])
  d = np.column_stack([rng.normal(0, 0.1, 400), rng.normal(0, 0.1, 400), -np.ones(400)])
  rs, rc = trace(oracle, sph, o, d), trace(oracle, con, o, d)
  assert np.array_equal(rs['n_segments'], rc['n_segments'])
  assert np.array_equal(rs['hits']['group'], rc['hits']['group']) and len(rs['hits']['group']) > 400
  assert np.abs(rs['hits']['points']-rc['hits']['points']).max() < 1e-11
  assert np.abs(rs['hits']['directions']-rc['hits']['directions']).max() < 1e-12


def test_parabolic_lens_surface_and_trim(oracle):
  'a plano-parabolic lens (conicoid + flat + rim): Snell at the curved face, rim and aperture respected'
  F, R, n = 20.0, 6.0, 1.5
  sag = R*R/(4*F)
  b = SceneBuilder()
  g = b.add_group('L', 'L', optical_type='Lens', refractive_index=n)
  faces = prim.conic_dish(1/(2*F), -1.0, R) + [
    prim._face(prim.plane_surface((0, 0, sag), (1, 0, 0), (0, 1, 0)), [prim._circle_loop(R)], shell_key=None)]
  for f in faces:
    f.shell_key = 1
  b.add_shape(g, faces, np.eye(4))
  a = b.add_group('A', 'A', optical_type='Absorber', record_hits=True)
  b.add_shape(a, prim.disc(100.0), prim.translation(0, 0, 60))
  scene = b.build()
  o = np.array([[1.0, 0.5, -10], [5.9, 0, -10], [6.1, 0, -10], [0, 0, -10]])
  d = np.tile([0, 0, 1.0], (4, 1))
  r = trace(oracle, scene, o, d)
  assert list(r['n_segments']) == [3, 3, 1, 3]            # the third ray misses the aperture and goes to the screen
  h = r['hits']
  P = h['points'][(h['ray_index'] == 0) & (h['bounce'] == 0)][0]
  assert abs(P[2]-(1+0.25)/(4*F)) < 1e-13
  # Snell at the first surface: n1 sin(i) = n2 sin(t) with the normal of the paraboloid
  nrm = np.array([P[0]/(2*F), P[1]/(2*F), -1.0]); nrm /= np.linalg.norm(nrm)
  din = np.array([0, 0, 1.0]); dout = h['directions'][(h['ray_index'] == 0) & (h['bounce'] == 1)][0]
  assert abs(np.linalg.norm(np.cross(din, nrm)) - n*np.linalg.norm(np.cross(dout, nrm))) < 1e-12
  # on-axis ray goes straight
  assert np.abs(h['points'][(h['ray_index'] == 3) & (h['bounce'] == 2)][0]-[0, 0, 60]).max() < 1e-12


def test_even_asphere_terms_known_answers(oracle):
  '''
  ODW_SEG_ASPHERE: sag = conic + a4 rho^4 + a6 rho^6.  Axis-parallel rays meet the surface at their own rho (z = sag(rho)
  exactly), oblique rays satisfy the implicit equation, the normal is the analytic gradient, and with all coefficients
  zero nothing changes against the plain conicoid.
  '''
  c, k, poly, R = 1/30.0, -0.6, [2.0e-5, -3.0e-8], 12.0
  from freecad.optics_design_workbench_b200.scene_export.scene import conic_sag
  asph = dish_and_screen(prim.conic_dish(c, k, R, poly=poly), screen_z=80.0)
  assert int(asph.faces[0]['seg_count']) == 1 and int(asph.segs[asph.faces[0]['seg_first']]['kind']) == sc_mod.SEG_ASPHERE
  o, d = grid_rays(R)
  r = trace(oracle, asph, o, d)
  h = r['hits']
  P1 = h['points'][h['bounce'] == 0]
  rho = np.hypot(o[:, 0], o[:, 1])
  assert len(P1) == len(o) and np.abs(P1[:, 2]-conic_sag(c, k, rho, poly)).max() < 1e-13
  # reflected direction from the analytic normal S'(rho) e_r - z
  u = rho*rho
  q = np.sqrt(1-(1+k)*c*c*u)
  slope = c*rho/q + 4*poly[0]*rho**3 + 6*poly[1]*rho**5
  n = np.column_stack([slope*o[:, 0]/rho, slope*o[:, 1]/rho, -np.ones_like(rho)])
  n /= np.linalg.norm(n, axis=1)[:, None]
  expect = d-2*np.sum(d*n, axis=1)[:, None]*n
  D2 = h['directions'][h['bounce'] == 1]
  assert np.abs(D2-expect).max() < 1e-12
  # oblique rays: the hit satisfies z = sag(rho) and lies inside the aperture
  rng = np.random.default_rng(4)
  oo = np.column_stack([rng.uniform(-9, 9, 500), rng.uniform(-9, 9, 500), np.full(500, 40.0)])
  dd = np.column_stack([rng.normal(0, 0.15, 500), rng.normal(0, 0.15, 500), -np.ones(500)])
  r = trace(oracle, asph, oo, dd, max_intersections=1)
  P = r['hits']['points'][r['hits']['group'] == 0]
  assert len(P) > 300
  rr = np.hypot(P[:, 0], P[:, 1])
  assert rr.max() < R+1e-6 and np.abs(P[:, 2]-conic_sag(c, k, rr, poly)).max() < 1e-12
  # a departure of 2e-5 rho^4 is visible: the plain conic puts the same rays elsewhere
  plain = dish_and_screen(prim.conic_dish(c, k, R), screen_z=80.0)
  r0 = trace(oracle, plain, oo, dd, max_intersections=1)
  assert np.abs(r0['hits']['points'][:50]-r['hits']['points'][:50]).max() > 1e-3
  # all-zero coefficients: no auxiliary record, identical results
  zero = dish_and_screen(prim.conic_dish(c, k, R, poly=[0.0, 0.0]), screen_z=80.0)
  assert int(zero.faces[0]['seg_count']) == 0
  rz = trace(oracle, zero, oo, dd, max_intersections=1)
  assert np.array_equal(rz['hits']['points'], r0['hits']['points'])


def test_asphere_corrects_spherical_aberration(oracle):
  '''
  A plano-convex singlet whose curved face is the hyperboloid with k = -n^2 focuses a collimated beam entering through the
  flat side to one point (the Cartesian-oval result for this configuration); the spherical face of the same curvature does not.
  Checks refraction through a conicoid face end to end.
  '''
  n, Rc, ap, thick = 1.5, 20.0, 8.0, 6.0
  def singlet(k):
    b = SceneBuilder()
    g = b.add_group('L', 'L', optical_type='Lens', refractive_index=n)
    # curved face bulging towards +z: the conicoid frame is flipped (vertex at z = thick, opening towards -z)
    faces = prim.conic_dish(1/Rc, k, ap) + [prim._face(prim.plane_surface((0, 0, 0), (1, 0, 0), (0, 1, 0)), [prim._circle_loop(ap)], reversed_=True, shell_key=None)]
    flip = prim.translation(0, 0, thick) @ prim.rotation((1, 0, 0), np.pi)
    faces[0].transform = flip
    for f in faces:
      f.shell_key = 1
    b.add_shape(g, faces, np.eye(4))
    a = b.add_group('A', 'A', optical_type='Absorber', record_hits=True)
    b.add_shape(a, prim.disc(50.0), prim.translation(0, 0, thick + Rc/(n-1)))       # paraxial focus: f = R/(n-1) behind the vertex
    return b.build()
  rho = np.linspace(0.5, 6.5, 13)
  o = np.column_stack([rho, np.zeros_like(rho), np.full_like(rho, -5.0)])
  d = np.tile([0, 0, 1.0], (len(rho), 1))
  spot = {}
  for k in (0.0, -n*n):
    r = trace(oracle, singlet(k), o, d)
    h = r['hits']
    assert np.all(r['n_segments'] == 3)
    spot[k] = np.abs(h['points'][h['group'] == 1][:, 0])
  assert spot[-n*n].max() < 1e-9            # stigmatic
  assert spot[0.0].max() > 0.05             # the sphere's marginal rays miss the paraxial focus by far more


def test_conicoid_validation():
  b = SceneBuilder()
  g = b.add_group('M', 'M', optical_type='Mirror')
  # beyond the equator of an oblate ellipsoid the sag formula has no value
  with pytest.raises(ValueError, match='equator'):
    b.add_shape(g, prim.conic_dish(0.1, 3.0, 20.0), np.eye(4))
  with pytest.raises(Exception):
    b.add_shape(g, prim.conic_dish(0.0, -1.0, 20.0), np.eye(4))


@pytest.mark.gpu
def test_gpu_conicoid_parity_with_oracle(gpu_engine, oracle):
  'paraboloid + ellipsoid + hyperboloid mirrors, a parabolic lens and a screen: same sequences, points to 1e-9 mm'
  b = SceneBuilder()
  m = b.add_group('M', 'M', optical_type='Mirror', reflectivity=0.9)
  b.add_shape(m, prim.conic_dish(1/25.0, -1.0, 20.0, poly=[1.0e-6]), np.eye(4))
  b.add_shape(m, prim.conic_dish(30.0/18.0**2, -(1-18.0**2/30.0**2), 15.0), prim.translation(45, 0, 0))
  b.add_shape(m, prim.conic_dish(10.0/(26.0**2-100.0), -(2.6**2), 15.0, inner_radius=2.0), prim.translation(-45, 0, 0) @ prim.rotation((1, 0, 0), 0.2))
  l = b.add_group('L', 'L', optical_type='Lens', refractive_index=1.5)
  # the lens carries even-asphere terms on top of its paraboloid (ODW_SEG_ASPHERE): Newton from the conic crossing
  lens = prim.conic_dish(1/40.0, -1.0, 8.0, poly=[-1.5e-5, 4.0e-8]) + [prim._face(prim.plane_surface((0, 0, 0.8-1.5e-5*8.0**4+4.0e-8*8.0**6), (1, 0, 0), (0, 1, 0)), [prim._circle_loop(8.0)], shell_key=None)]
  for f in lens:
    f.shell_key = 1
  b.add_shape(l, lens, prim.translation(0, 0, 30))
  a = b.add_group('A', 'A', optical_type='Absorber', record_hits=True)
  b.add_shape(a, prim.disc(300.0), prim.translation(0, 0, 90))
  scene = b.build()
  assert sum(int(f['kind']) == sc_mod.SURF_CONICOID for f in scene.faces) == 4
  assert sum(int(k) == sc_mod.SEG_ASPHERE for k in scene.segs['kind']) == 2
  rng = np.random.default_rng(21)
  n = 20000
  o = np.column_stack([rng.uniform(-60, 60, n), rng.uniform(-18, 18, n), np.full(n, 80.0)])
  d = np.column_stack([rng.normal(0, 0.08, n), rng.normal(0, 0.08, n), -np.ones(n)])
  cfg = _abi.CfgArgs(max_ray_length=500.0, dist_tol=1e-6, max_intersections=20, record_all_hits=True, hit_capacity=25*n)
  with gpu_engine.scene(scene).trace_rays(cfg, o, d) as res:
    gh = res.hits(sort=True)
  ref = oracle.trace_rays(scene, cfg, o, d, hit_capacity=25*n, threads=0)
  oh = ref['hits']
  assert len(gh['face_id']) == len(oh['face_id']) > n
  assert np.array_equal(gh['face_id'], oh['face_id']) and np.array_equal(gh['ray_index'], oh['ray_index'])
  assert np.array_equal(gh['is_entering'], oh['is_entering'])
  assert np.abs(gh['points']-oh['points']).max() < 1e-9
  assert np.abs(gh['directions']-oh['directions']).max() < 1e-9
  assert np.abs(gh['powers']-oh['powers']).max() < 1e-14
  assert {int(scene.faces[f]['kind']) for f in np.unique(gh['face_id'])} >= {1, 6}


# ---- aspheres the way FreeCAD users model them: a spline through points of the lens formula, revolved ----------------
ASPHERE = dict(c=1/25.0, k=-0.8, poly=[2e-5, -3e-8])


def spline_dish(n_points=81, **kw):
  sag = lambda r: sc_mod.conic_sag(ASPHERE['c'], ASPHERE['k'], r, ASPHERE['poly'])
  return prim.revolved_spline_dish(sag, 8.0, n_points=n_points, **kw)


def test_revolved_spline_meridian_is_fitted_to_the_even_asphere_form(oracle):
  '''
  A surface of revolution with a B-spline generatrix (BRep surface type 7) becomes a closed-form conicoid + even-asphere
  terms when the fit reproduces the meridian within ASPHERE_FIT_TOLERANCE (1e-7 mm); the stated residual is the error
  bound of the face.  Rays reflected by it land where the exact asphere sends them, within that bound's effect.
  '''
  faces = spline_dish()
  fitted = sc_mod._revolved_asphere(faces[0].surface)
  assert fitted is not None and fitted.fit_residual < 1e-8              # the cubic spline through 81 points is that close to the formula
  assert abs(fitted.c-ASPHERE['c']) < 1e-8 and abs(fitted.k-ASPHERE['k']) < 1e-3
  scene = dish_and_screen(faces, screen_z=12.0, screen_r=50.0)
  f = scene.faces[0]
  assert int(f['kind']) == sc_mod.SURF_CONICOID and int(f['trim_kind']) == sc_mod.TRIM_UVBOX
  assert abs(f['uv_max'][1]-8.0) < 1e-9 and f['uv_min'][1] == 0.0      # the trim is in rho now, not in the curve parameter
  assert int(scene.segs[int(f['seg_first'])]['kind']) == sc_mod.SEG_ASPHERE
  exact = dish_and_screen(prim.conic_dish(ASPHERE['c'], ASPHERE['k'], 8.0, poly=ASPHERE['poly']), screen_z=12.0, screen_r=50.0)
  o, d = grid_rays(8.0)
  a, b = trace(oracle, scene, o, d), trace(oracle, exact, o, d)
  assert np.array_equal(a['n_segments'], b['n_segments']) and np.array_equal(a['hits']['group'], b['hits']['group'])
  assert np.abs(a['hits']['points']-b['hits']['points']).max() < 1e-6   # residual 1e-9 mm in sag -> ~1e-8 in slope -> < 1e-6 mm on the screen
  assert np.abs(a['hits']['directions']-b['hits']['directions']).max() < 1e-7


def test_a_meridian_that_is_no_asphere_is_not_fitted():
  'a wavy meridian leaves a residual above the tolerance: no closed form, the face goes to the tessellation as before'
  wavy = prim.revolved_spline_dish(lambda r: 0.02*r*r + 0.01*np.sin(3*r), 8.0, n_points=81)
  assert sc_mod._revolved_asphere(wavy[0].surface) is None
  b = SceneBuilder()
  g = b.add_group('M', 'M', optical_type='Mirror')
  b.add_shape(g, wavy, np.eye(4))
  scene = b.build()
  assert len(scene.faces) > 100 and (scene.faces['kind'] == sc_mod.SURF_PLANE).all()      # meshed: planar triangles
  coarse = spline_dish(n_points=6)                                       # a spline through 6 points is 1e-4 mm off the formula ...
  fitted = sc_mod._revolved_asphere(coarse[0].surface)
  assert fitted is None or fitted.fit_residual <= sc_mod.ASPHERE_FIT_TOLERANCE   # ... and is only accepted if the fit really reproduces IT


@pytest.mark.gpu
def test_gpu_fitted_asphere_parity_with_oracle(gpu_engine, oracle):
  scene = dish_and_screen(spline_dish(), screen_z=12.0, screen_r=50.0)
  o, d = grid_rays(8.0, n=25)
  cfg = _abi.CfgArgs(max_ray_length=1000.0, dist_tol=1e-6, max_intersections=20, record_all_hits=True)
  want = oracle.trace_rays(scene, cfg, o, d)
  ds = gpu_engine.scene(scene)
  with ds.trace_rays(cfg, o, d) as res:
    got = res.hits(sort=True)
  ds.close()
  assert np.array_equal(got['face_id'], want['hits']['face_id']) and np.array_equal(got['ray_index'], want['hits']['ray_index'])
  assert np.abs(got['points']-want['hits']['points']).max() < 1e-8
  assert np.abs(got['directions']-want['hits']['directions']).max() < 1e-9
