'''
Scene export: BRep reader, FCStd importer, placement resolver, fixtures.
Tests marked `reference` read the reference's FCStd files from /root/reference (build container only).
'''
import glob
import os
import zipfile

import numpy as np
import pytest

from freecad.optics_design_workbench_b200.scene_export import brep, fcstd, primitives as prim, scene as sc
from freecad.optics_design_workbench_b200.simulation.setup import PreparedSimulation

REF = '/root/reference'


@pytest.mark.reference
def test_every_brep_in_the_reference_tree_parses():
  n = 0
  for f in sorted(glob.glob(REF+'/**/*.FCStd', recursive=True)):
    z = zipfile.ZipFile(f)
    for m in z.namelist():
      if m.endswith('.brp'):
        t = z.read(m).decode('ascii', 'replace')
        if not t.strip():
          continue
        faces = brep.read_brep(t).faces()
        for fi in faces:
          assert fi.loops and all(fi.loops), (f, m, fi.surface.kind)      # every face has a boundary in (u,v)
          assert abs(abs(np.linalg.det(fi.transform[:3, :3]))-1) < 1e-9
        n += 1
  assert n > 500


@pytest.mark.reference
def test_benchmark_scene_geometry_matches_survey_appendix_b():
  sim = PreparedSimulation.from_fcstd(REF+'/benchmark/lensesAndMirrors.FCStd')
  s = sim.scene
  assert s.summary()['census'] == {'cylinder/uvbox': 2, 'plane/loops': 2, 'plane/uvbox': 18, 'sphere/uvbox': 2, 'torus/none': 1}
  assert s.group_names == ['OpticalMirrorGroup', 'OpticalLensGroup', 'OpticalMirrorGroup001', 'OpticalAbsorberGroup']
  assert [int(g['optical_type']) for g in s.groups] == [sc.OPT_MIRROR, sc.OPT_LENS, sc.OPT_MIRROR, sc.OPT_ABSORBER]
  assert [int(g['record_hits']) for g in s.groups] == [0, 0, 0, 1]
  spheres = s.faces[s.faces['kind'] == sc.SURF_SPHERE]
  assert np.allclose(spheres['origin'], [[-28, 0, 32], [-30, 0, 32]]) and np.allclose(spheres['p0'], 5)
  tor = s.faces[s.faces['kind'] == sc.SURF_TORUS][0]
  assert np.allclose(tor['origin'], [-70, 0, 67]) and (tor['p0'], tor['p1']) == (10.0, 2.0)
  assert sim.settings['MaxRayLength'] == pytest.approx(460.1823980554802)
  assert sim.settings['DistanceTolerance'] == 1e-6 and not sim.settings['SequentialMode']
  rec = sim.source_records[0]
  assert rec['PowerDensity'] == 'exp(-theta**2/(1e-2)**2)' and rec['ThetaResolutionNumericMode'] == '1e5'


@pytest.mark.reference
def test_sequential_lists_and_link_array():
  sim = PreparedSimulation.from_fcstd(REF+'/benchmark/lensesAndMirrorsSequential.FCStd')
  assert list(sim.scene.seq_offsets) == [0, 1, 2, 3, 4, 5] and list(sim.scene.seq_groups) == [0, 1, 1, 2, 3]
  sim = PreparedSimulation.from_fcstd(REF+'/benchmark/hugeArray.FCStd')
  f = sim.scene.faces
  assert len(f) == 1500 and np.all(f['kind'] == sc.SURF_SPHERE) and np.all(f['trim_kind'] == sc.TRIM_NONE)
  for g, z0 in ((0, 26), (1, 0), (2, 51)):         # mirror, lens, absorber arrays
    c = f['origin'][f['group'] == g]
    assert len(c) == 500
    assert np.allclose(np.unique(c[:, 0]), -23+5*np.arange(10)) and np.allclose(np.unique(c[:, 2]), z0+5*np.arange(5))


@pytest.mark.reference
def test_global_placements_of_nested_links():
  '8 placements of ShiftedCube, reference test/22-global-placement/z-freecad-placements.py:42-51'
  doc = fcstd.FCStdDocument(REF+'/test/22-global-placement/main.FCStd')
  cube = [o for o in doc.objects.values() if o.Label == 'ShiftedCube' or o.Name == 'ShiftedCube']
  assert cube, [o.Label for o in doc.objects.values()]
  got = sorted(tuple(np.round(m[:3, 3], 9)) for m, _ in doc.global_placements(cube[0]))
  expect = sorted([(0, 0, -100), (3, 3, -100), (3, 0, -100), (3, -27, -100), (3, -27, -100), (3, 3, -97),
                   (0, 0, -100), (0, -30, -100)])
  assert got == [tuple(float(x) for x in e) for e in expect]


@pytest.mark.reference
def test_fixtures_are_up_to_date():
  'tests/golden/scenes/*.npz equal a fresh export of the reference scenes'
  here = os.path.dirname(__file__)
  for name in ('minimal', 'lensesAndMirrors', 'lensesAndMirrorsSequential', 'hugeArray'):
    fresh = PreparedSimulation.from_fcstd(f'{REF}/benchmark/{name}.FCStd')
    fix = PreparedSimulation.from_fixture(os.path.join(here, 'golden', 'scenes', name+'.npz'))
    assert fresh.scene.faces.tobytes() == fix.scene.faces.tobytes()
    assert fresh.scene.segs.tobytes() == fix.scene.segs.tobytes()
    assert fresh.scene.groups.tobytes() == fix.scene.groups.tobytes()
    assert fresh.settings['MaxRayLength'] == fix.settings['MaxRayLength']


def test_fixture_round_trip(tmp_path, sims):
  sim = sims('lensesAndMirrorsSequential')
  p = tmp_path/'x.npz'
  sim.save_fixture(p)
  again = PreparedSimulation.from_fixture(p)
  assert again.scene.faces.tobytes() == sim.scene.faces.tobytes()
  assert list(again.scene.seq_groups) == list(sim.scene.seq_groups)
  assert again.settings['SequentialMode'] and np.isinf(again.settings['EndAfterRays'])
  assert np.array_equal(again.source_records[0]['gpM'], sim.source_records[0]['gpM'])


def test_trim_classification_of_primitives():
  b = sc.SceneBuilder()
  g = b.add_group('G', 'G', 'Lens')
  for faces in (prim.box(1, 2, 3), prim.sphere(2), prim.torus(5, 1), prim.cylinder(1, 4), prim.cone(2, 1, 3),
                prim.plano_convex_lens(10, 2, 0.3), prim.disc(3)):
    b.add_shape(g, faces, np.eye(4))
  s = b.build()
  census = s.summary()['census']
  assert census == {'cone/uvbox': 1, 'cylinder/uvbox': 2, 'plane/loops': 6, 'plane/uvbox': 6, 'sphere/none': 1,
                    'sphere/uvbox': 1, 'torus/none': 1}
  assert len(s.shells) == 7
  # outward normals: nsign * (xdir x ydir) of every box face points away from the box centre
  for f in s.faces[:6]:
    n = f['nsign']*np.cross(f['xdir'], f['ydir'])
    centre = sc.eval_face(f, (f['uv_min'][0]+f['uv_max'][0])/2, (f['uv_min'][1]+f['uv_max'][1])/2)
    assert np.dot(n, centre-np.array([0.5, 1, 1.5])) > 0


def test_even_odd_trim_with_arcs_and_lines():
  segs = [(sc.SEG_ARC, [0, 0, 2.0, 0.0, 2*np.pi]), (sc.SEG_ARC, [0.5, 0, 0.5, 0.0, 2*np.pi])]   # annulus with an off-centre hole
  assert sc.point_in_segs(segs, 1.5, 0.2) and not sc.point_in_segs(segs, 0.5, 0.1) and not sc.point_in_segs(segs, 2.5, 0)
  half = [(sc.SEG_ARC, [0, 0, 1.0, 0.0, np.pi]), (sc.SEG_LINE, [-1, 0, 1, 0, 0])]                # upper half disc
  assert sc.point_in_segs(half, 0.2, 0.5) and not sc.point_in_segs(half, 0.2, -0.5)


def test_face_aabb_contains_the_face():
  rng = np.random.default_rng(0)
  b = sc.SceneBuilder()
  g = b.add_group('G', 'G', 'Lens')
  m = prim.translation(3, -2, 7) @ prim.rotation((1, 2, 3), 0.7)
  for faces in (prim.sphere(2), prim.torus(5, 1), prim.cylinder(1, 4), prim.cone(2, 1, 3), prim.plano_convex_lens(10, 2, 0.3)):
    b.add_shape(g, faces, m)
  for f in b.build().faces:
    u = rng.uniform(f['uv_min'][0], f['uv_max'][0], 2000)
    v = rng.uniform(f['uv_min'][1], f['uv_max'][1], 2000)
    p = sc.eval_face(f, u, v)
    assert np.all(p >= f['aabb_min']-1e-12) and np.all(p <= f['aabb_max']+1e-12)


@pytest.mark.reference
def test_live_export_from_freecad_like_objects():
  '''
  scene_export/freecad_live.py with stand-ins for FreeCAD objects: Shape.exportBrepToString() returns the stored BRep of
  benchmark/minimal.FCStd's absorber box, the placement matrices are what allCoordinateTransformMatrices would give.
  '''
  import types, zipfile
  from freecad.optics_design_workbench_b200.scene_export import freecad_live, fcstd
  path = os.path.join(REF, 'benchmark', 'minimal.FCStd')
  doc = fcstd.FCStdDocument(path)
  group = doc.optical_groups()[0]
  child = doc.objects[group.get('ElementList')[0]]
  text = zipfile.ZipFile(path).read(child.get('Shape')).decode()
  gp = doc.global_placements(group)[0][0]
  class M:                                                  # FreeCAD.Matrix look-alike
    def __init__(self, a):
      for r in range(4):
        for c in range(4):
          setattr(self, f'A{r+1}{c+1}', float(a[r, c]))
  pM = group.placement
  obj = types.SimpleNamespace(Name=group.Name, Label=group.Label, OpticalType='Absorber', RecordHits=True,
                              Shape=types.SimpleNamespace(exportBrepToString=lambda: text))
  # group.Shape = pM * child shape in the reference; the stored child BRep stands in for it, so pMi = identity here
  scene, info = freecad_live.build_scene([obj], lambda g: [[M(gp), M(np.linalg.inv(gp)), M(np.eye(4)), M(np.eye(4))]])
  ref_scene, _ = fcstd.build_scene(doc)
  assert len(scene.faces) == len(ref_scene.faces) == 6 and not info['skipped']
  np.testing.assert_allclose(np.sort(scene.faces['aabb_min'], axis=0), np.sort(ref_scene.faces['aabb_min'], axis=0), atol=1e-12)
  np.testing.assert_allclose(np.sort(scene.faces['aabb_max'], axis=0), np.sort(ref_scene.faces['aabb_max'], axis=0), atol=1e-12)
  assert scene.group_names == ref_scene.group_names and int(scene.groups[0]['record_hits']) == 1
