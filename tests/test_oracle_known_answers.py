'''
Known-answer tests of the CPU oracle (oracle/odw_oracle.c): hand-derived results for the reference's
benchmark scenes and closed-form optics on procedural scenes.  These are what
"pins" the trace semantics here, because the reference's own tests hold no per-ray golden vector and
FreeCAD/OCC cannot run in this image (oracle header: "parity unpinned" for per-ray sequences).
'''
import numpy as np
import pytest

from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.scene_export import primitives as prim
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder

S2 = np.sqrt(0.5)


def ray_dir(theta, phi):
  'point-source convention (reference point_source.py:429-432, quirk Q6: phi=0 points to -y)'
  return np.array([np.sin(theta)*np.sin(phi), -np.sin(theta)*np.cos(phi), np.cos(theta)])


def trace(oracle, scene, origins, dirs, **cfg):
  kw = dict(max_ray_length=1000.0, dist_tol=1e-6, max_intersections=100, record_all_hits=True)
  kw.update(cfg)
  return oracle.trace_rays(scene, _abi.CfgArgs(**kw), np.atleast_2d(origins), np.atleast_2d(dirs))


# ------------------------------------------------------------------------------------------
# benchmark scenes

def test_minimal_known_answer(oracle, sims):
  'hit (15 tan(theta) sin(phi), -15 tan(theta) cos(phi), 15), 1 segment, power 1, entering'
  sim = sims('minimal')
  th = np.array([0.0, 0.01, 0.05, 0.2, 0.3])
  ph = np.array([0.0, 1.0, 2.5, 4.0, 6.0])
  d = np.array([ray_dir(t, p) for t, p in zip(th, ph)])
  r = oracle.trace_rays(sim.scene, sim.cfg(), np.zeros_like(d), d)
  h = r['hits']
  expect = np.stack([15*np.tan(th)*np.sin(ph), -15*np.tan(th)*np.cos(ph), np.full_like(th, 15)], axis=-1)
  assert np.array_equal(r['n_segments'], np.ones(5, dtype=np.int32))
  assert np.abs(h['points']-expect).max() < 1e-12
  assert np.abs(h['directions']-d).max() == 0
  assert np.all(h['powers'] == 1) and np.all(h['is_entering'] == 1)
  assert np.all(r['final_powers'] == 0)          # absorbed


def test_lenses_and_mirrors_on_axis(oracle, sims):
  'SURVEY.md Appendix B: 7 segments through mirror, lens, lens (link), mirror, torus hole, absorber box'
  sim = sims('lensesAndMirrors')
  r = oracle.trace_rays(sim.scene, sim.cfg(record_all_hits=True), [[0, 0, 0]], [[0, 0, 1]])
  h = r['hits']
  x6 = 5*S2*2 + 5*S2*2 - 81          # 5.071 + 7.071 - 81 with exact roots
  expect = np.array([[0, 0, 32], [-23, 0, 32], [-24, 0, 32], [-34, 0, 32], [-35, 0, 32],
                     [-68.85786437626905, 0, 32], [-68.85786437626905, 0, 73]])
  assert r['n_segments'][0] == 7
  assert np.abs(h['points']-expect).max() < 1e-9
  names = [sim.scene.group_names[g] for g in h['group']]
  assert names == ['OpticalMirrorGroup', 'OpticalLensGroup', 'OpticalLensGroup', 'OpticalLensGroup',
                   'OpticalLensGroup', 'OpticalMirrorGroup001', 'OpticalAbsorberGroup']
  assert list(h['is_entering']) == [1, 1, 0, 1, 0, 1, 1]
  assert np.abs(h['directions'][1]-[-1, 0, 0]).max() < 1e-12 and np.abs(h['directions'][6]-[0, 0, 1]).max() < 1e-12
  # only the absorber records by default
  r = oracle.trace_rays(sim.scene, sim.cfg(), [[0, 0, 0]], [[0, 0, 1]])
  assert r['counts']['hits'] == 1 and sim.scene.group_names[r['hits']['group'][0]] == 'OpticalAbsorberGroup'


def test_sequential_scene_object_sequence(oracle, sims):
  'every non-edge ray: Mirror, Lens x4, Mirror001, Absorber (SURVEY.md Appendix B)'
  sim = sims('lensesAndMirrorsSequential')
  assert sim.settings['SequentialMode']
  rng = np.random.default_rng(3)
  th, ph = np.abs(rng.normal(0, 7e-3, 200)), rng.uniform(0, 2*np.pi, 200)
  d = np.array([ray_dir(t, p) for t, p in zip(th, ph)])
  r = oracle.trace_rays(sim.scene, sim.cfg(record_all_hits=True), np.zeros_like(d), d)
  assert np.all(r['n_segments'] == 7)
  seq = r['hits']['group'].reshape(200, 7)
  assert np.all(seq == np.array([0, 1, 1, 1, 1, 2, 3]))
  # identical geometry traced non-sequentially gives the same hits
  sim2 = sims('lensesAndMirrors')
  r2 = oracle.trace_rays(sim2.scene, sim2.cfg(record_all_hits=True), np.zeros_like(d), d)
  assert np.abs(r2['hits']['points']-r['hits']['points']).max() < 1e-12


def test_huge_array_first_hit_is_nearest_sphere(oracle, sims):
  sim = sims('hugeArray')
  sa = sim.source_args(0)
  s = oracle.sample_mc(sa, 7, 0, 300)
  r = oracle.trace_rays(sim.scene, sim.cfg(record_all_hits=True), s['origins'], s['directions'])
  h = r['hits']
  first = h['bounce'] == 0
  centres = sim.scene.faces['origin']
  for ray, P, fid in zip(h['ray_index'][first], h['points'][first], h['face_id'][first]):
    o, d = s['origins'][ray], s['directions'][ray]
    # brute-force nearest unit sphere
    w = o-centres
    b = w @ d
    disc = b*b-(np.einsum('ij,ij->i', w, w)-1.0)
    t = np.where(disc > 0, -b-np.sqrt(np.maximum(disc, 0)), np.inf)
    t[t <= 1e-6] = np.inf
    assert fid == int(np.argmin(t))
    assert abs(np.linalg.norm(P-centres[fid])-1.0) < 1e-10


# ------------------------------------------------------------------------------------------
# optics on procedural scenes

def one_group_scene(faces, transform=np.eye(4), **group):
  b = SceneBuilder()
  g = b.add_group('G', 'G', **group)
  b.add_shape(g, faces, transform)
  return b


def test_mirror_reflection_and_reflectivity(oracle):
  b = one_group_scene(prim.box(10, 10, 1), prim.translation(-5, -5, 10), optical_type='Mirror', reflectivity=0.5)
  sc = b.build()
  d = np.array([np.sin(0.3), 0, np.cos(0.3)])
  r = trace(oracle, sc, [0, 0, 0], d, max_ray_length=50.0)
  assert r['n_segments'][0] == 2                       # hit + escaping segment
  assert np.abs(r['hits']['points'][0]-[10*np.tan(0.3), 0, 10]).max() < 1e-12
  assert r['final_powers'][0] == 0.5
  expect_end = r['hits']['points'][0] + 50.0*np.array([np.sin(0.3), 0, -np.cos(0.3)])
  assert np.abs(r['final_points'][0]-expect_end).max() < 1e-9
  assert r['counts']['escaped'] == 1


def test_snell_slab_and_displacement(oracle):
  'plane-parallel plate n=1.5: exit direction equals entry direction, lateral shift matches the textbook formula'
  n, t, a = 1.5, 4.0, 0.6
  b = one_group_scene(prim.box(100, 100, t), prim.translation(-50, -50, 10), optical_type='Lens', refractive_index=n)
  b2 = b.add_group('D', 'D', optical_type='Absorber', record_hits=True)
  b.add_shape(b2, prim.rectangle(200, 200), prim.translation(-100, -100, 30))
  sc = b.build()
  d = np.array([np.sin(a), 0, np.cos(a)])
  r = trace(oracle, sc, [0, 0, 0], d)
  h = r['hits']
  assert r['n_segments'][0] == 3 and list(h['is_entering'][:2]) == [1, 0]
  inside = h['directions'][1]
  assert abs(np.sin(a)/n - inside[0]) < 1e-12          # Snell
  assert np.abs(h['directions'][2]-d).max() < 1e-12     # parallel exit
  shift = t*np.sin(a)*(1-np.cos(a)/np.sqrt(n*n-np.sin(a)**2))
  x_no_plate = 30*np.tan(a)
  assert abs((x_no_plate-h['points'][2][0])*np.cos(a)-shift) < 1e-10


def test_total_internal_reflection_keeps_medium(oracle):
  'ray born inside a glass block (medium unknown -> n1 = 1 on the first exit, quirk noted at ray.py:185-189)'
  b = one_group_scene(prim.box(10, 10, 10), optical_type='Lens', refractive_index=1.5)
  sc = b.build()
  # enter through the bottom face at a steep angle, then hit a side wall beyond the critical angle
  a = 1.2
  d = np.array([np.sin(a), 0, np.cos(a)])
  r = trace(oracle, sc, [2, 5, -1], d, max_ray_length=100.0)
  h = r['hits']
  # first hit enters (bottom), second is the +x wall: incidence angle inside = 90deg - asin(sin(a)/1.5) > critical
  assert h['is_entering'][0] == 1 and h['is_entering'][1] == 0
  din = h['directions'][1]
  dout = h['directions'][2]
  assert abs(dout[0]+din[0]) < 1e-12 and abs(dout[2]-din[2]) < 1e-12     # mirrored at the x wall
  assert r['n_segments'][0] >= 4


def test_paraxial_focus_of_plano_convex_lens(oracle):
  'thin-ish plano-convex lens, f = R/(n-1) measured from the curved vertex for rays entering the flat side'
  R, n = 50.0, 1.5
  b = one_group_scene(prim.plano_convex_lens(R, 5.0, 0.5), prim.translation(0, 0, 20), optical_type='Lens', refractive_index=n)
  sc = b.build()
  hs = np.array([0.05, 0.1, 0.2])
  o = np.stack([hs, np.zeros(3), np.zeros(3)], axis=-1)
  d = np.tile([0, 0, 1.0], (3, 1))
  r = trace(oracle, sc, o, d, max_ray_length=500.0)
  h = r['hits']
  assert np.all(r['n_segments'] == 3)
  exit_pts = h['points'][h['bounce'] == 1]
  # extrapolate the escaping segment to the axis
  ends = r['final_points']
  dirs = ends-exit_pts
  tz = -exit_pts[:, 0]/dirs[:, 0]
  zf = exit_pts[:, 2]+tz*dirs[:, 2]
  vertex = 20 + 0.5 + (R-np.sqrt(R*R-25.0))
  assert np.abs(zf-(vertex+R/(n-1))).max() < 0.05      # spherical aberration is ~h^2/R
  assert abs(zf[0]-(vertex+R/(n-1))) < 5e-3


def test_absorber_vacuum_and_record_flags(oracle):
  'Vacuum detector: two hits per traversal (enter + exit), no change of direction/power (ray.py:276-277)'
  b = SceneBuilder()
  v = b.add_group('V', 'V', optical_type='Vacuum', record_hits=True)
  b.add_shape(v, prim.box(10, 10, 2), prim.translation(-5, -5, 5))
  a = b.add_group('A', 'A', optical_type='Absorber', record_hits=True)
  b.add_shape(a, prim.sphere(1.0), prim.translation(0, 0, 20))
  sc = b.build()
  r = oracle.trace_rays(sc, _abi.CfgArgs(max_ray_length=100.0), [[0.2, 0.1, 0]], [[0, 0, 1]])
  h = r['hits']
  assert r['n_segments'][0] == 3 and len(h['powers']) == 3
  assert list(h['group']) == [0, 0, 1] and list(h['is_entering']) == [1, 0, 1]
  assert np.abs(h['points'][:, 2]-[5, 7, 20-np.sqrt(1-0.05)]).max() < 1e-12
  assert r['final_powers'][0] == 0


def test_max_intersections_and_escape(oracle):
  'two facing mirrors: the ray bounces until maxIntersections stops it (ray.py:96-98)'
  b = SceneBuilder()
  m = b.add_group('M', 'M', optical_type='Mirror')
  b.add_shape(m, prim.box(10, 10, 1), prim.translation(-5, -5, 10))
  b.add_shape(m, prim.box(10, 10, 1), prim.translation(-5, -5, -11))
  sc = b.build()
  r = trace(oracle, sc, [0, 0, 0], [0, 0, 1], max_intersections=13)
  assert r['n_segments'][0] == 13 and r['counts']['depth_terminated'] == 1 and r['counts']['escaped'] == 0
  r = trace(oracle, sc, [0, 0, 0], [1, 0, 0], max_ray_length=77.0)
  assert r['n_segments'][0] == 1 and r['counts']['escaped'] == 1 and r['counts']['hits'] == 0
  assert np.abs(r['final_points'][0]-[77, 0, 0]).max() < 1e-12


def test_sequential_mode_filters_groups(oracle):
  'sequence [B], [A]: the nearer A is invisible for the first segment; nothing is hittable after the sequence ends'
  b = SceneBuilder()
  a = b.add_group('A', 'A', optical_type='Vacuum', record_hits=True)
  b.add_shape(a, prim.rectangle(10, 10), prim.translation(-5, -5, 5))
  bb = b.add_group('B', 'B', optical_type='Mirror', record_hits=True)
  b.add_shape(bb, prim.rectangle(10, 10), prim.translation(-5, -5, 10))
  sc_seq = b.build(sequence=[[bb], [a]])
  r = trace(oracle, sc_seq, [0, 0, 0], [0, 0, 1], sequential=True, max_ray_length=100.0)
  assert list(r['hits']['group']) == [bb, a] and r['n_segments'][0] == 3
  r = trace(oracle, sc_seq, [0, 0, 0], [0, 0, 1], sequential=False, max_ray_length=100.0)
  assert list(r['hits']['group'])[:2] == [a, bb]


def test_ignored_groups(oracle):
  b = SceneBuilder()
  a = b.add_group('A', 'A', optical_type='Absorber', record_hits=True)
  b.add_shape(a, prim.rectangle(10, 10), prim.translation(-5, -5, 5))
  c = b.add_group('C', 'C', optical_type='Absorber', record_hits=True)
  b.add_shape(c, prim.rectangle(10, 10), prim.translation(-5, -5, 9))
  sc = b.build()
  cfg = _abi.CfgArgs(max_ray_length=100.0)
  assert list(oracle.trace_rays(sc, cfg, [[0, 0, 0]], [[0, 0, 1]])['hits']['group']) == [a]
  assert list(oracle.trace_rays(sc, cfg, [[0, 0, 0]], [[0, 0, 1]], ignored=[a])['hits']['group']) == [c]


def test_prefers_other_group_at_coincident_faces(oracle):
  'two lenses touching at z=10: leaving A and entering B coincide; the hit that is not the current medium wins (ray.py:445-452)'
  b = SceneBuilder()
  a = b.add_group('A', 'A', optical_type='Lens', refractive_index=1.5)
  b.add_shape(a, prim.box(10, 10, 5), prim.translation(-5, -5, 5))
  c = b.add_group('B', 'B', optical_type='Lens', refractive_index=1.7)
  b.add_shape(c, prim.box(10, 10, 5), prim.translation(-5, -5, 10))
  sc = b.build()
  r = trace(oracle, sc, [0, 0, 0], [np.sin(0.2), 0, np.cos(0.2)], max_ray_length=100.0)
  h = r['hits']
  assert list(h['group']) == [a, c, c] and list(h['is_entering']) == [1, 1, 0]
  # refraction at the A->B interface uses n1 = n(A), n2 = n(B)
  assert abs(h['directions'][2][0]-np.sin(0.2)/1.7) < 1e-12


def test_cylinder_cone_torus_intersections(oracle):
  b = SceneBuilder()
  g = b.add_group('A', 'A', optical_type='Absorber', record_hits=True)
  b.add_shape(g, prim.cylinder(2.0, 10.0), prim.translation(0, 0, 0))
  b.add_shape(g, prim.cone(3.0, 1.0, 4.0), prim.translation(20, 0, 0))
  b.add_shape(g, prim.torus(10.0, 2.0), prim.translation(60, 0, 0))
  sc = b.build()
  o = np.array([[-10, 0.5, 3], [20-10, 0.3, 1.0], [60, 0, 10], [60+10.3, 0.2, 10], [60-30, 0, 0.5]])
  d = np.array([[1, 0, 0], [1, 0, 0], [0, 0, -1], [0, 0, -1], [1, 0, 0.0]])
  r = oracle.trace_rays(sc, _abi.CfgArgs(max_ray_length=200.0), o, d)
  h = r['hits']
  P = {int(k): p for k, p in zip(h['ray_index'], h['points'])}
  assert abs(P[0][0]+np.sqrt(4-0.25)) < 1e-12                               # cylinder wall
  rad = 3.0-0.5*1.0                                                       # cone radius at z=1
  assert abs(P[1][0]-(20-np.sqrt(rad*rad-0.09))) < 1e-12
  assert 2 not in P                                                       # straight through the torus hole
  x, y = 0.3, 0.2
  rho = np.hypot(10+x, y)
  assert abs(P[3][2]-np.sqrt(4-(rho-10)**2)) < 1e-12                      # top of the tube
  assert abs(P[4][0]-(60-10-np.sqrt(4-0.25))) < 1e-12                     # outer equator side


def test_grating_littrow_like_first_order(oracle):
  'reflection grating at normal incidence: sin(theta_m) = m*lambda/d (Ludwig 1970 form in ray.py:497-539)'
  lpm, wl = 600.0, 500.0
  b = SceneBuilder()
  g = b.add_group('G', 'G', optical_type='Grating', grating_type='Reflection', grating_lines_per_mm=lpm,
                  grating_order=1, grating_orientation=(0, 1, 0))
  b.add_shape(g, prim.box(10, 10, 1), prim.translation(-5, -5, 10))
  sc = b.build()
  r = oracle.trace_rays(sc, _abi.CfgArgs(max_ray_length=100.0, record_all_hits=True), [[0, 0, 0]], [[0, 0, 1]], wavelength=wl)
  out = r['final_points'][0]-r['hits']['points'][0]
  out /= np.linalg.norm(out)
  s = 1*(wl*1e-6)/(1.0/lpm)          # m*lambda/d with lambda, d in mm
  assert out[2] < 0                   # reflected
  # GratingLinesOrientation g=(0,1,0) is the normal of the planes that cut the rulings: rulings run along
  # P = g x n = x, dispersion happens along D = n x P = y
  assert abs(abs(out[1])-s) < 1e-9 and abs(out[0]) < 1e-12


def test_finite_absorption_length_is_multiplicative(oracle):
  'deliberate divergence Q1 (SURVEY.md): Beer-Lambert multiplies; the reference line ray.py:125 cannot run'
  b = SceneBuilder()
  g = b.add_group('L', 'L', optical_type='Lens', refractive_index=1.0, absorption_length=5.0)
  b.add_shape(g, prim.box(10, 10, 10), prim.translation(-5, -5, 10))
  sc = b.build()
  r = trace(oracle, sc, [0, 0, 0], [0, 0, 1], max_ray_length=100.0)
  assert abs(r['final_powers'][0]-np.exp(-10/5.0)) < 1e-15
