'''
Tessellation path of the scene export: faces whose surface has no closed form here (B-spline, Bezier, surfaces of
revolution / extrusion, offset surfaces) become planar triangles with a stated deflection.

Every triangle is written as an ordinary plane face of the flat scene (FACE_DTYPE: plane through the three vertices,
trim = the three edges as line segments in the plane's (u, v)), so the engine needs nothing new: scenes with more than
64 faces go through the SAH BVH (csrc/odw_api.cu BvhBuilder) and the exact fp64 plane + trim test, and an emitting
face of a surface source is sampled triangle by triangle with the area weights of the mesh.  What the reference asks
OpenCASCADE for on such faces (line/surface intersection ray.py:411, normalAt :465, valueAt / derivative1At
surface_source.py:282-316) is therefore approximated: positions to within `deflection`, normals piecewise constant.

Method: regular (u, v) grid over the bounding box of the trim loops, refined (doubling) until the distance between the
surface at the cell centres and the bilinear interpolant of the cell corners is below `deflection` (or `max_grid` is
reached); cells whose centre lies inside the trim region (even-odd rule over the face's pcurves) are kept and split
into two triangles.  The trim boundary is therefore reproduced to within one cell.
'''

import numpy as np

from . import scene as sc

DEFAULT_DEFLECTION = 1e-3     # mm
MAX_GRID = 128                # finest (u, v) grid per face: at most 2*MAX_GRID**2 triangles


def points_in_segs(segs, U, V):
  'vectorised even-odd test (see scene.point_in_segs): segs = list of (kind, a); U, V arrays -> bool array'
  U, V = np.asarray(U, dtype=float), np.asarray(V, dtype=float)
  crossings = np.zeros(U.shape, dtype=np.int64)
  lines = np.array([a[:4] for k, a in segs if k == sc.SEG_LINE], dtype=float).reshape(-1, 4)
  for chunk in np.array_split(lines, max(1, len(lines)//512)) if len(lines) else []:
    u0, v0, u1, v1 = (chunk[:, i].reshape((-1,)+(1,)*U.ndim) for i in range(4))
    cond = (v0 > V) != (v1 > V)
    with np.errstate(divide='ignore', invalid='ignore'):
      ux = u0 + (V-v0)*(u1-u0)/(v1-v0)
    crossings += (cond & (ux > U)).sum(axis=0)
  for k, a in segs:
    if k != sc.SEG_ARC:
      continue
    cu, cv, r, a0, span = a[:5]
    dv = V-cv
    ok = np.abs(dv) < r
    h = np.sqrt(np.where(ok, r*r-dv*dv, 0.0))
    for ux in (cu-h, cu+h):
      rel = (np.arctan2(dv, ux-cu)-a0) % sc.TWO_PI
      crossings += (ok & (ux > U) & (rel <= span)).astype(np.int64)
  return (crossings & 1) == 1


def _uv_window(fi, segs):
  if segs:
    return sc._segs_bbox(segs)
  s = getattr(fi.surface, 'spline', None)
  if s is not None:
    return np.array([s['uknots'][0], s['vknots'][0]]), np.array([s['uknots'][-1], s['vknots'][-1]])
  raise sc.UnsupportedGeometry('face without boundary on a surface without natural bounds')


def tessellate(fi, deflection=DEFAULT_DEFLECTION, max_grid=None, min_grid=8):
  '''
  brep.FaceInstance -> (triangles [m, 3, 3] in the coordinates of the stored shape, grid size, achieved deflection).
  Triangle vertex order follows dS/du x dS/dv.
  '''
  surf = fi.surface
  segs = []
  for loop in fi.loops:
    for curve, first, last in loop:
      sc._curve_to_segs(curve, first, last, out=segs)
  lo, hi = _uv_window(fi, segs)
  max_grid = MAX_GRID if max_grid is None else max_grid
  n = min_grid
  while True:
    u, v = np.linspace(lo[0], hi[0], n+1), np.linspace(lo[1], hi[1], n+1)
    U, V = np.meshgrid(u, v, indexing='ij')
    P = surf.eval(U, V)                                                    # [n+1, n+1, 3]
    uc, vc = (u[1:]+u[:-1])/2, (v[1:]+v[:-1])/2
    Uc, Vc = np.meshgrid(uc, vc, indexing='ij')
    inside = points_in_segs(segs, Uc, Vc) if segs else np.ones(Uc.shape, dtype=bool)
    corners = (P[:-1, :-1]+P[1:, :-1]+P[:-1, 1:]+P[1:, 1:])/4
    if inside.any():
      Pc = surf.eval(Uc[inside], Vc[inside])
      err = float(np.linalg.norm(Pc-corners[inside], axis=-1).max())
    else:
      err = np.inf
    if (err <= deflection and inside.any()) or n >= max_grid:
      break
    n *= 2
  if not inside.any():
    raise sc.UnsupportedGeometry('tessellation found no grid cell inside the trim region')
  i, j = np.nonzero(inside)
  p00, p10, p11, p01 = P[i, j], P[i+1, j], P[i+1, j+1], P[i, j+1]
  tris = np.concatenate([np.stack([p00, p10, p11], axis=1), np.stack([p00, p11, p01], axis=1)], axis=0)
  # drop degenerate triangles (collapsed parameter lines at poles)
  e1, e2 = tris[:, 1]-tris[:, 0], tris[:, 2]-tris[:, 0]
  area2 = np.linalg.norm(np.cross(e1, e2), axis=1)
  return tris[area2 > 1e-14], n, err


def triangle_faces(fi, transform, group, shell, face_id, segs_out, deflection=DEFAULT_DEFLECTION, max_grid=None):
  '''
  FaceInstance -> list of FACE_DTYPE rows (one plane face per triangle, world frame), all carrying `face_id`;
  their trim segments are appended to segs_out.
  '''
  tris, n, err = tessellate(fi, deflection, max_grid)
  R, T = sc._rigid(transform @ fi.transform)
  tris = tris @ R.T + T
  out = []
  sign = -1 if fi.reversed else 1
  for a, b, c in tris:
    e1, e2 = b-a, c-a
    nrm = np.cross(e1, e2)
    nl = np.linalg.norm(nrm)
    x = e1/np.linalg.norm(e1)
    z = nrm/nl
    y = np.cross(z, x)
    f = np.zeros((), dtype=sc.FACE_DTYPE)
    f['origin'], f['xdir'], f['ydir'], f['zdir'] = a, x, y, z
    f['kind'], f['nsign'] = sc.SURF_PLANE, sign
    f['group'], f['shell'], f['face_id'] = group, shell, face_id
    q1, q2 = (float(e1@x), 0.0), (float(e2@x), float(e2@y))
    loop = [(sc.SEG_LINE, [0.0, 0.0, q1[0], q1[1], 0.0]), (sc.SEG_LINE, [q1[0], q1[1], q2[0], q2[1], 0.0]),
            (sc.SEG_LINE, [q2[0], q2[1], 0.0, 0.0, 0.0])]
    f['trim_kind'] = sc.TRIM_LOOPS
    f['uv_min'] = [min(0.0, q1[0], q2[0]), min(0.0, q2[1])]
    f['uv_max'] = [max(0.0, q1[0], q2[0]), max(0.0, q2[1])]
    f['seg_first'], f['seg_count'] = len(segs_out), 3
    segs_out.extend(loop)
    pts = np.stack([a, b, c])
    f['aabb_min'], f['aabb_max'] = pts.min(axis=0), pts.max(axis=0)
    out.append(f)
  return out, dict(grid=n, deflection=err, triangles=len(out))
