'''
The OpenCASCADE stand-ins of tests/occ_stub.py (they answer for OCC when the reference's own findNearestIntersection
is executed for the golden, tests/golden/make_traceray_golden.py) checked against answers known in closed form.
'''
import math

import numpy as np

import freecad_stub
import occ_stub
from freecad_stub import Vector


def surface(kind, p0=0.0, p1=0.0, origin=(0, 0, 0)):
  return occ_stub.Surface(kind, origin, (1, 0, 0), (0, 1, 0), (0, 0, 1), p0, p1)


def test_line_crossings_of_the_elementary_surfaces():
  s = surface(occ_stub.SPHERE, 5.0, origin=(0, 0, 20))
  assert np.allclose(s.line_parameters((3, 0, 0), (0, 0, 1)), [16.0, 24.0], atol=1e-12)          # 3-4-5
  assert s.line_parameters((6, 0, 0), (0, 0, 1)) == []
  c = surface(occ_stub.CYLINDER, 2.0)
  assert np.allclose(c.line_parameters((-5, 1, 3), (1, 0, 0)), [5-math.sqrt(3), 5+math.sqrt(3)], atol=1e-12)
  assert c.line_parameters((1, 0, 0), (0, 0, 1)) == []                                           # parallel to the axis
  k = surface(occ_stub.CONE, 1.0, math.pi/4)                                                     # radius 1 + z
  assert np.allclose(k.line_parameters((-10, 0, 2), (1, 0, 0)), [7.0, 13.0], atol=1e-12)
  t = surface(occ_stub.TORUS, 10.0, 2.0)
  assert np.allclose(t.line_parameters((-20, 0, 0), (1, 0, 0)), [8, 12, 28, 32], atol=1e-10)
  assert np.allclose(t.line_parameters((10, 0, -5), (0, 0, 1)), [3, 7], atol=1e-10)
  assert t.line_parameters((0, 0, -5), (0, 0, 1)) == []                                          # through the hole
  p = surface(occ_stub.PLANE, origin=(0, 0, 7))
  assert np.allclose(p.line_parameters((1, 2, 3), (0, 0.6, 0.8)), [5.0], atol=1e-12)
  assert p.line_parameters((1, 2, 3), (1, 0, 0)) == []


def test_parameters_follow_elslib():
  s = surface(occ_stub.SPHERE, 2.0)
  u, v = s.parameter((0, -2, 0))
  assert abs(u-1.5*math.pi) < 1e-15 and abs(v) < 1e-15                                          # u in [0, 2 pi)
  t = surface(occ_stub.TORUS, 10.0, 2.0)
  u, v = t.parameter((8, 0, 0))
  assert abs(u) < 1e-15 and abs(v-math.pi) < 1e-15
  assert np.allclose(t.value(u, v), (8, 0, 0), atol=1e-14)
  k = surface(occ_stub.CONE, 1.0, math.pi/4)
  u, v = k.parameter((3, 0, 2))
  assert np.allclose(k.value(u, v), (3, 0, 2), atol=1e-14)
  n = k.normal_geom(u, v)
  assert np.allclose(n, (math.sqrt(0.5), 0, -math.sqrt(0.5)), atol=1e-15)


def test_bounding_box_line_test_is_two_sided():
  b = occ_stub.BoundBox((0, 0, 10), (1, 1, 11))
  assert b.intersect(Vector(0.5, 0.5, 20), Vector(0, 0, 1))               # the box lies BEHIND the start: still cut (quirk Q4)
  assert not b.intersect(Vector(2, 0.5, 0), Vector(0, 0, 1))
  assert b.isInside(Vector(1, 1, 11)) and not b.isInside(Vector(1, 1, 11.0001))
  assert tuple(b.closestPoint(Vector(5, 0.5, 0))) == (1, 0.5, 10)
  b.enlarge(0.5)
  assert b.isInside(Vector(1.4, -0.4, 9.6))


def face(kind, trim, uv_min=(0, 0), uv_max=(0, 0), p0=0.0, p1=0.0, segs=()):
  from freecad.optics_design_workbench_b200.scene_export.scene import FACE_DTYPE, SEG_DTYPE
  rec = np.zeros((), dtype=FACE_DTYPE)
  rec['xdir'], rec['ydir'], rec['zdir'] = (1, 0, 0), (0, 1, 0), (0, 0, 1)
  rec['kind'], rec['trim_kind'], rec['nsign'], rec['p0'], rec['p1'] = kind, trim, 1, p0, p1
  rec['uv_min'], rec['uv_max'] = uv_min, uv_max
  table = np.zeros(len(segs), dtype=SEG_DTYPE)
  for row, (k, a) in zip(table, segs):
    row['kind'], row['a'] = k, a
  rec['seg_first'], rec['seg_count'] = 0, len(segs)
  return occ_stub.Face(rec, table, 0)


def test_distance_to_trimmed_faces():
  rect = face(occ_stub.PLANE, occ_stub.TRIM_UVBOX, (0, 0), (4, 2))
  assert rect.distance(np.array([1.0, 1.0, 0.0])) == 0.0
  assert abs(rect.distance(np.array([1.0, 1.0, 0.25]))-0.25) < 1e-15
  assert abs(rect.distance(np.array([7.0, 6.0, 0.0]))-5.0) < 1e-9         # nearest point is the corner (4, 2)
  assert abs(rect.distance(np.array([2.0, -0.5, 0.0]))-0.5) < 1e-9
  disc = face(occ_stub.PLANE, occ_stub.TRIM_LOOPS, segs=[(occ_stub.SEG_ARC, (0, 0, 2.0, 0.0, 2*math.pi))])
  assert disc.distance(np.array([1.9, 0.0, 0.0])) == 0.0
  assert abs(disc.distance(np.array([0.0, 3.0, 0.0]))-1.0) < 1e-9
  band = face(occ_stub.CYLINDER, occ_stub.TRIM_UVBOX, (0, 0), (2*math.pi, 10), p0=2.0)
  assert band.distance(np.array([0.0, -2.0, 5.0])) < 1e-15
  assert abs(band.distance(np.array([2.0, 0.0, 10.5]))-0.5) < 1e-9
  cap = face(occ_stub.SPHERE, occ_stub.TRIM_UVBOX, (0, math.asin(0.6)), (2*math.pi, math.pi/2), p0=5.0)
  assert cap.distance(np.array([0.0, 3.0, 4.0])) < 1e-12                  # inside the cap (z > 3)
  assert abs(cap.distance(np.array([5.0, 0.0, 0.0]))-math.hypot(1.0, 3.0)) < 1e-9   # equator point -> rim circle (rho 4, z 3)
