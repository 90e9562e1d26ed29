'''
Host side of the surface light source (Lambertian-style emission from faces).

Mirrors SurfaceSourceProxy of the reference (reference freecad_elements/surface_source.py):
  _generateRays 'true'      :418-466,522-555   which faces emit (ActiveSurfaces x placements of the part), weighted by
                                               face area; theta from the power density WITHOUT sin(theta), phi uniform
  _drawRandomPositionOnFace :390-410           area-uniform point on the trimmed face  -> done in the kernel
  _makeRay                  :85-111            direction from (theta, phi), face normal and tangent -> done in the kernel
This module turns the emitting faces into the `emit_*` arrays of odw_source_desc: world-frame face records,
their trim segments and the cumulative area weights.  Face areas come in closed form for untrimmed /
(u,v)-box-trimmed elementary faces and from an exact-per-row scanline integral of the even-odd trim region otherwise.
'''

import numpy as np

from ..scene_export import scene as sc

TWO_PI = 2*np.pi


def _row_metric(f, v):
  'area element dA/(du dv) of the parametrisation on the row v'
  k = int(f['kind'])
  if k == sc.SURF_PLANE:
    return 1.0
  if k == sc.SURF_CYLINDER:
    return abs(float(f['p0']))
  if k == sc.SURF_CONE:
    return abs(float(f['p0']) + v*np.sin(float(f['p1'])))
  if k == sc.SURF_SPHERE:
    return float(f['p0'])**2*np.cos(v)
  return (float(f['p0']) + float(f['p1'])*np.cos(v))*abs(float(f['p1']))


def _window(f):
  k, trim = int(f['kind']), int(f['trim_kind'])
  u0, u1, v0, v1 = float(f['uv_min'][0]), float(f['uv_max'][0]), float(f['uv_min'][1]), float(f['uv_max'][1])
  if trim == sc.TRIM_NONE:
    u0, u1 = 0.0, TWO_PI
    if k == sc.SURF_SPHERE:
      v0, v1 = -np.pi/2, np.pi/2
    if k == sc.SURF_TORUS:
      v0, v1 = 0.0, TWO_PI
  return u0, u1, v0, v1


def _v_integral(f, v0, v1):
  'closed form of the integral of _row_metric over [v0, v1]'
  k = int(f['kind'])
  if k in (sc.SURF_PLANE, sc.SURF_CYLINDER):
    return _row_metric(f, 0.0)*(v1-v0)
  if k == sc.SURF_SPHERE:
    return float(f['p0'])**2*(np.sin(v1)-np.sin(v0))
  if k == sc.SURF_TORUS:
    R, r = float(f['p0']), abs(float(f['p1']))
    return r*(R*(v1-v0) + float(f['p1'])*(np.sin(v1)-np.sin(v0)))
  # cone: |p0 + v sin a|, possibly changing sign inside the window
  p0, sa = float(f['p0']), np.sin(float(f['p1']))
  F = lambda v: p0*v + 0.5*sa*v*v
  if sa == 0:
    return abs(p0)*(v1-v0)
  vz = -p0/sa
  if v0 < vz < v1:
    return abs(F(vz)-F(v0)) + abs(F(v1)-F(vz))
  return abs(F(v1)-F(v0))


def _scanline_area(f, segs, rows=4096):
  '''
  Area of an even-odd trimmed face: for every row v (mid-points of `rows` strips) the crossings of the row with all
  trim segments are computed exactly, sorted, and the inside intervals' u-lengths summed; times the row metric.
  '''
  u0, u1, v0, v1 = _window(f)
  kinds = np.array([int(s['kind']) for s in segs])
  A = np.array([s['a'] for s in segs], dtype=np.float64).reshape(-1, 5)
  lines, arcs = A[kinds == sc.SEG_LINE], A[kinds == sc.SEG_ARC]
  vs = v0 + (np.arange(rows)+0.5)*(v1-v0)/rows
  total = 0.0
  for v in vs:
    xs = []
    if len(lines):
      la, lb = lines[:, 1], lines[:, 3]
      m = (la > v) != (lb > v)
      if m.any():
        l = lines[m]
        xs.append(l[:, 0] + (v-l[:, 1])*(l[:, 2]-l[:, 0])/(l[:, 3]-l[:, 1]))
    for cu, cv, r, a0, span in arcs:
      dv = v-cv
      if abs(dv) < r:
        h = np.sqrt(r*r-dv*dv)
        for ux in (cu-h, cu+h):
          if (np.arctan2(dv, ux-cu)-a0) % TWO_PI <= span:
            xs.append(np.array([ux]))
    if not xs:
      continue
    x = np.sort(np.concatenate(xs))
    if len(x) % 2:
      x = x[:-1]
    total += (x[1::2]-x[0::2]).sum()*_row_metric(f, v)
  return total*(v1-v0)/rows


def face_area(f, segs):
  'area of FACE_DTYPE row f (its trim segments in `segs`, SEG_DTYPE rows indexed by seg_first/seg_count)'
  trim = int(f['trim_kind'])
  u0, u1, v0, v1 = _window(f)
  if trim in (sc.TRIM_NONE, sc.TRIM_UVBOX):
    return (u1-u0)*_v_integral(f, v0, v1)
  mine = segs[int(f['seg_first']):int(f['seg_first'])+int(f['seg_count'])]
  if int(f['kind']) == sc.SURF_PLANE and len(mine) == 3 and all(int(m['kind']) == sc.SEG_LINE for m in mine):
    (x0, y0, x1, y1), (_, _, x2, y2) = mine[0]['a'][:4], mine[1]['a'][:4]
    return 0.5*abs((x1-x0)*(y2-y0)-(x2-x0)*(y1-y0))                        # a triangle
  if (int(f['kind']) == sc.SURF_PLANE and len(mine) == 1 and int(mine[0]['kind']) == sc.SEG_ARC
          and mine[0]['a'][4] >= TWO_PI-1e-12):
    return np.pi*float(mine[0]['a'][2])**2            # a disc
  return _scanline_area(f, mine)


class EmittingFaces:
  'faces + trim segments + areas of one surface source, world frame'
  def __init__(self, faces, segs):
    self.faces = np.ascontiguousarray(faces, dtype=sc.FACE_DTYPE)
    self.segs = np.ascontiguousarray(segs, dtype=sc.SEG_DTYPE)
    self.areas = np.array([face_area(f, self.segs) for f in self.faces], dtype=np.float64)
    if len(self.areas) and not np.all(self.areas > 0):
      raise ValueError(f'emitting face with non-positive area: {self.areas}')

  @property
  def cdf(self):
    'cumulative area weights (surface_source.py:465-466), last entry exactly 1'
    c = np.cumsum(self.areas)/self.areas.sum()
    c[-1] = 1.0
    return c


def emitting_faces_from_instances(selections):
  '''
  selections: list of (brep.FaceInstance list, 4x4 world transform) — one entry per (selected faces of a part) x
  (placement of the part); what the reference collects in allFacesAndPlacements (:437-458).
  '''
  faces, segs = [], []
  for face_instances, transform in selections:
    for fi in face_instances:
      try:
        faces.append(sc.face_record(fi, transform, 0, 0, len(faces), segs))
      except sc.UnsupportedGeometry:
        # B-spline and other free-form emitters: emit from the triangles of the tessellated face
        from ..scene_export import tessellate
        tris, _info = tessellate.triangle_faces(fi, transform, 0, 0, len(faces), segs)
        faces.extend(tris)
  seg_arr = np.zeros(len(segs), dtype=sc.SEG_DTYPE)
  for i, (kind, a) in enumerate(segs):
    seg_arr[i]['kind'] = kind
    seg_arr[i]['a'] = a
  return EmittingFaces(np.array(faces, dtype=sc.FACE_DTYPE) if faces else np.zeros(0, dtype=sc.FACE_DTYPE), seg_arr)
