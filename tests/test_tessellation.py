'''
Tessellation path (scene_export/tessellate.py): free-form faces -> planar triangles with a stated deflection, traced as
ordinary plane faces through the BVH.  Checked on a surface whose closed form is known (a sphere pushed through the
tessellator), against the analytic face; and on the reference's B-spline scenes (marker `reference`).
'''
import types

import numpy as np
import pytest

from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.freecad_elements import surface_source
from freecad.optics_design_workbench_b200.scene_export import primitives as prim, scene as sc, tessellate
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder

R = 5.0


def sphere_instance():
  return prim.sphere(R)[0]          # FaceInstance of the analytic sphere; the tessellator only uses surface.eval


def test_tessellated_sphere_meets_the_stated_deflection():
  segs = []
  faces, info = tessellate.triangle_faces(sphere_instance(), prim.translation(1, 2, 3), 0, 0, 7, segs, deflection=2e-3)
  assert info['deflection'] <= 2e-3 and info['triangles'] == len(faces) > 1000
  c = np.array([1.0, 2.0, 3.0])
  for f in faces[::97]:
    assert int(f['kind']) == sc.SURF_PLANE and int(f['trim_kind']) == sc.TRIM_LOOPS and int(f['face_id']) == 7
    tri = segs[int(f['seg_first']):int(f['seg_first'])+3]
    verts = [f['origin'] + a[0]*f['xdir'] + a[1]*f['ydir'] for _, a in tri]
    for v in verts:
      assert abs(np.linalg.norm(v-c)-R) < 1e-9                            # vertices on the surface
    centroid = np.mean(verts, axis=0)
    assert 0 <= R-np.linalg.norm(centroid-c) < 4e-3                       # chord sag of the order of the deflection
    assert (f['zdir']*f['nsign']) @ (centroid-c) > 0                      # outward orientation kept
  seg_arr = np.zeros(len(segs), dtype=sc.SEG_DTYPE)
  for i, (k, a) in enumerate(segs):
    seg_arr[i]['kind'], seg_arr[i]['a'] = k, a
  area = sum(surface_source.face_area(f, seg_arr) for f in faces)
  assert abs(area-4*np.pi*R*R)/(4*np.pi*R*R) < 2e-3


def scenes():
  'the same absorbing sphere once analytic, once tessellated'
  out = []
  for meshed in (False, True):
    b = SceneBuilder()
    g = b.add_group('Abs', 'Abs', optical_type='Absorber', record_hits=True)
    if meshed:
      first = len(b.faces)
      tris, _ = tessellate.triangle_faces(sphere_instance(), prim.translation(0, 0, 30), g, 0, first, b.segs, deflection=1e-3)
      b.faces.extend(tris)
      sh = np.zeros((), dtype=sc.SHELL_DTYPE)
      sh['aabb_min'] = np.min([f['aabb_min'] for f in tris], axis=0); sh['aabb_max'] = np.max([f['aabb_max'] for f in tris], axis=0)
      sh['face_first'], sh['face_count'], sh['group'] = first, len(tris), g
      b.shells.append(sh)
    else:
      b.add_shape(g, prim.sphere(R), prim.translation(0, 0, 30))
    out.append(b.build())
  return out


def fan(n=4000, spread=0.12, seed=3):
  rng = np.random.default_rng(seed)
  th, ph = np.abs(rng.normal(0, spread, n)), rng.uniform(0, 2*np.pi, n)
  return np.zeros((n, 3)), np.stack([np.sin(th)*np.sin(ph), -np.sin(th)*np.cos(ph), np.cos(th)], axis=-1)


def test_hits_on_the_mesh_agree_with_the_analytic_sphere_within_the_deflection(oracle):
  analytic, meshed = scenes()
  o_, d_ = fan()
  cfg = _abi.CfgArgs(max_ray_length=200.0, record_all_hits=True)
  a = oracle.trace_rays(analytic, cfg, o_, d_, threads=0)
  m = oracle.trace_rays(meshed, cfg, o_, d_, threads=0)
  ia, im = a['hits']['ray_index'], m['hits']['ray_index']
  both = np.intersect1d(ia, im)
  assert len(both) > 0.98*len(ia) > 1000                                  # grazing rays may miss the inscribed mesh
  pa = a['hits']['points'][np.searchsorted(ia, both)]
  pm = m['hits']['points'][np.searchsorted(im, both)]
  dist = np.linalg.norm(pa-pm, axis=1)
  # the mesh hit lies within ~2 deflections of the true surface (cell-centre criterion + the diagonal split) ...
  sag = np.abs(np.linalg.norm(pm-[0, 0, 30], axis=1)-R)
  assert sag.max() < 2.5e-3
  # ... and along the ray that sag is amplified by 1/cos(incidence): 95 % of the rays within 5 deflections
  assert np.quantile(dist, 0.95) < 5e-3 and np.median(dist) < 1e-3


@pytest.mark.gpu
def test_gpu_traces_the_mesh_like_the_oracle(gpu_engine, oracle):
  _, meshed = scenes()
  assert len(meshed.faces) > 64                                           # BVH + wavefront path
  o_, d_ = fan(20000)
  cfg = _abi.CfgArgs(max_ray_length=200.0, record_all_hits=True, hit_capacity=40000)
  with gpu_engine.scene(meshed).trace_rays(cfg, o_, d_) as res:
    gc, gh = res.counts, res.hits(sort=True)
  o = oracle.trace_rays(meshed, cfg, o_, d_, hit_capacity=40000, threads=0)
  # rays through a shared edge of two triangles may pick either (both within distTol): same ray set, same positions
  np.testing.assert_array_equal(gh['ray_index'], o['hits']['ray_index'])
  np.testing.assert_allclose(gh['points'], o['hits']['points'], rtol=0, atol=1e-6)
  assert gc['hits'] == o['counts']['hits']


@pytest.mark.reference
def test_reference_lambert_source_scene_emits_from_bspline_faces(oracle, monkeypatch):
  'test/50-old-tests/lambert-source.FCStd: the emitters are a sphere and a scaled (B-spline) clone of it'
  from freecad.optics_design_workbench_b200.simulation.setup import prepare
  monkeypatch.setattr(tessellate, 'MAX_GRID', 16)                          # keep the CPU test quick
  sim = prepare('/root/reference/test/50-old-tests/lambert-source.FCStd')
  rec = sim.source_records[0]
  assert rec['emit_error'] is None and len(rec['emit'].faces) > 500
  sa = sim.source_args(0)
  s = oracle.sample_mc(sa, 1, 0, 20000)
  assert np.isfinite(s['origins']).all() and np.allclose(np.linalg.norm(s['directions'], axis=1), 1.0)
  r = oracle.trace_mc(sim.scene, sa, sim.cfg(), 1, 0, 5000, threads=0)
  assert r['counts']['segments'] >= 5000
