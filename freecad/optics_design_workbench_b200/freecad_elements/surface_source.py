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


# ------------------------------------------------------------------------------------------------------------------
# fan mode (reference surface_source.py:122-267 _makeSurfaceGrid, :469-517 the 'fans' branch of _generateRays):
# a deterministic, approximately equidistant grid of points on every emitting face, one ray along the face normal
# from each.  Host side: at most FanModeRayCount rays, traced as an explicit list (odw_trace_rays).

class FaceEvaluator:
  '''
  What _makeSurfaceGrid asks OpenCASCADE about a face, in closed form on a FACE_DTYPE row in the world frame:
  ParameterRange, valueAt, derivative1At, normalAt and "is this surface point on the trimmed face within distTol"
  (the reference's Part.Vertex(p).distToShape(face)[0] < distTol, :176-178).
  '''
  def __init__(self, f, segs):
    self.f = f
    self.segs = segs[int(f['seg_first']):int(f['seg_first'])+int(f['seg_count'])]
    self.kind, self.trim = int(f['kind']), int(f['trim_kind'])
    self.O, self.X, self.Y, self.Z = (np.asarray(f[k], dtype=np.float64) for k in ('origin', 'xdir', 'ydir', 'zdir'))
    self.p0, self.p1 = float(f['p0']), float(f['p1'])

  @property
  def area(self):
    return face_area(self.f, self.segs if self.trim == sc.TRIM_LOOPS else self.segs)

  def parameter_range(self):
    return _window(self.f)

  def value_at(self, u, v):
    return sc.eval_face(self.f, u, v).reshape(3)

  def derivative1_at(self, u, v):
    k = self.kind
    if k == sc.SURF_PLANE:
      return self.X.copy(), self.Y.copy()
    er = np.cos(u)*self.X + np.sin(u)*self.Y
    et = -np.sin(u)*self.X + np.cos(u)*self.Y
    if k == sc.SURF_CYLINDER:
      return self.p0*et, self.Z.copy()
    if k == sc.SURF_CONE:
      return (self.p0 + v*np.sin(self.p1))*et, np.sin(self.p1)*er + np.cos(self.p1)*self.Z
    if k == sc.SURF_SPHERE:
      return self.p0*np.cos(v)*et, self.p0*(-np.sin(v)*er + np.cos(v)*self.Z)
    return (self.p0 + self.p1*np.cos(v))*et, self.p1*(-np.sin(v)*er + np.cos(v)*self.Z)

  def normal_at(self, u, v):
    'unit normal pointing out of the solid (orientation-aware, like Face.normalAt)'
    k = self.kind
    er = np.cos(u)*self.X + np.sin(u)*self.Y
    if k == sc.SURF_PLANE:
      n = self.Z
    elif k == sc.SURF_CYLINDER:
      n = er
    elif k == sc.SURF_CONE:
      sgn = 1.0 if self.p0 + v*np.sin(self.p1) >= 0 else -1.0
      n = sgn*(np.cos(self.p1)*er - np.sin(self.p1)*self.Z)
    else:                                       # sphere and torus: same expression in their own (u, v)
      n = np.cos(v)*er + np.sin(v)*self.Z
    return float(self.f['nsign'])*n

  def _uv_boundary_distance(self, u, v):
    'distance of (u, v) to the trim loops, in parameter units'
    best = np.inf
    for s in self.segs:
      a = s['a']
      if int(s['kind']) == sc.SEG_LINE:
        p, q = np.array([a[0], a[1]]), np.array([a[2], a[3]])
        d = q-p
        t = 0.0 if not d.any() else min(1.0, max(0.0, float(np.dot([u-p[0], v-p[1]], d)/np.dot(d, d))))
        best = min(best, float(np.hypot(u-(p[0]+t*d[0]), v-(p[1]+t*d[1]))))
      else:
        cu, cv, r, a0, span = a[:5]
        ang = (np.arctan2(v-cv, u-cu)-a0) % TWO_PI
        if ang <= span:
          best = min(best, abs(float(np.hypot(u-cu, v-cv))-r))
        else:
          for e in (a0, a0+span):
            best = min(best, float(np.hypot(u-(cu+r*np.cos(e)), v-(cv+r*np.sin(e)))))
    return best

  def on_face(self, u, v, tol):
    if self.trim == sc.TRIM_NONE:
      return True
    du, dv = self.derivative1_at(u, v)
    lu, lv = max(np.linalg.norm(du), 1e-300), max(np.linalg.norm(dv), 1e-300)
    if self.trim == sc.TRIM_UVBOX:
      u0, u1, v0, v1 = _window(self.f)
      return bool(u0-tol/lu <= u <= u1+tol/lu and v0-tol/lv <= v <= v1+tol/lv)
    if sc.point_in_segs(self.segs, u, v):
      return True
    return bool(self._uv_boundary_distance(u, v)*max(lu, lv) < tol)


def make_surface_grid(face, total_grid_points, dist_tol, uniform_param=None, fill_factor=1, effective_sizes=None,
                      recursion_depth=0):
  '''
  _makeSurfaceGrid (surface_source.py:122-267) on a FaceEvaluator: five passes that refine (i) which parameter is laid
  out first, (ii) the effective lengths of the two parameter axes, (iii) the fraction of the rectangular (u, v) grid
  that lies on the face; rows of a "uniform" parametrisation are thinned by powers of two where they are short (the
  poles of a sphere).  Returns [((u, v), point, (du, dv))].  numpy's sum / mean / round are used where the reference's
  `from numpy import *` makes it use them.
  '''
  r = face.parameter_range()
  limits = dict(u=(r[0], r[1]), v=(r[2], r[3]))
  param_sizes = dict(u=r[1]-r[0], v=r[3]-r[2])
  if effective_sizes is None:
    effective_sizes = param_sizes
  order = 'uv' if effective_sizes['u'] >= effective_sizes['v'] else 'vu'
  if uniform_param is not None:
    order = uniform_param + {'u': 'v', 'v': 'u'}[uniform_param]
  P1, P2 = [np.linspace(limits[p][0], limits[p][1],
                        max(5, 1+int(2*np.round(np.sqrt(effective_sizes[p]/effective_sizes[q]*total_grid_points/fill_factor)/2))))
            for p, q in zip(order, reversed(order))]
  p1_step, p2_step = P1[1]-P1[0], P2[1]-P2[0]
  uv = (lambda a, b: (a, b)) if order == 'uv' else (lambda a, b: (b, a))
  points = [[face.value_at(*uv(p1, p2)) for p2 in P2] for p1 in P1]
  valid = [[face.on_face(*uv(p1, p2), dist_tol) for p2 in P2] for p1 in P1]
  deriv = [[uv(*face.derivative1_at(*uv(p1, p2))) if valid[i][j] else (None, None) for j, p2 in enumerate(P2)]
           for i, p1 in enumerate(P1)]
  length = lambda d: None if d is None else float(np.sqrt(d[0]*d[0]+d[1]*d[1]+d[2]*d[2]))
  d1 = [[length(d[0]) for d in row] for row in deriv]
  d2 = [[length(d[1]) for d in row] for row in deriv]
  area = [[a*b if a is not None and b is not None else None for a, b in zip(r1, r2)] for r1, r2 in zip(d1, d2)]
  some = lambda A: [a for a in A if a is not None]
  mean_ = lambda A: np.mean(some(A)) if len(some(A)) else None
  max_ = lambda A: np.max(some(A)) if len(some(A)) else None
  sum_ = lambda A: np.sum(some(A))
  uniform = lambda rows: all(all(abs(d-avg)*p1_step*p2_step < dist_tol**2 for d in row if d is not None)
                             for avg, row in zip([mean_(row) for row in rows], rows))
  if uniform(area):
    uniform_param = order[0]
  elif uniform(list(zip(*area))):
    uniform_param = order[1]
  else:
    uniform_param = None
  eff1 = sum_([max_(row) for row in d1])*p1_step
  eff2 = sum_([max_(col) for col in zip(*d2)])*p2_step
  effective_sizes = {order[0]: eff1, order[1]: eff2}
  if uniform_param is not None:
    for i, row in enumerate(d2):
      row_len = max(1e-20, p2_step*sum_(row))
      keep_every = 2**np.round(np.log2(eff2/row_len))
      if keep_every > len(valid[i]):
        for j in range(1, len(valid[i])):
          valid[i][j] = False
      else:
        for j in range(len(valid[i])):
          if j % keep_every != 0:
            valid[i][j] = False
  count = sum(sum(1 if v else 0 for v in row) for row in valid)
  if recursion_depth < 4:
    return make_surface_grid(face, total_grid_points, dist_tol, uniform_param=uniform_param,
                             fill_factor=max(fill_factor/10, count/(len(P1)*len(P2))),
                             effective_sizes=effective_sizes, recursion_depth=recursion_depth+1)
  drops = [lambda i, j: False, lambda i, j: i % 2 == 0 or j % 2 == 0, lambda i, j: ((i+1)//2) % 2 == 0 or ((j+1)//2) % 2 == 0]
  while len(drops) and total_grid_points < 20 and count > total_grid_points:
    drop = drops.pop(0)
    valid = [[valid[i][j] and not drop(i, j) for j in range(len(P2))] for i in range(len(P1))]
    count = sum(sum(1 if v else 0 for v in row) for row in valid)
  return [(uv(p1, p2), points[i][j], uv(*deriv[i][j])) for i, p1 in enumerate(P1) for j, p2 in enumerate(P2) if valid[i][j]]


def _custom_round(x):
  'ray counts per face: 1, 4, 9 or any larger integer (surface_source.py:474-476)'
  return round(x) if x > 9 else [1, 4, 9][int(np.argmin(np.abs(x-np.array([1, 4, 9]))))]


def fan_ray_counts(weights, fan_mode_ray_count):
  '''
  which faces get rays and how many (surface_source.py:478-503): by area weight, but at least one per face; when that
  overshoots FanModeRayCount by more than 30 % a matching fraction of the faces is skipped.  Returns [(face index, rays)].
  The reference itself raises NameError on that branch (its warning text uses the undefined names `warnings` and
  `rayCount`, :485-488); the skipping rule below is its code taken literally, without the warning.
  '''
  total = sum(_custom_round(w*fan_mode_ray_count) for w in weights)
  skip_fraction = max(0, 1-fan_mode_ray_count/total)
  if skip_fraction <= 0.3:
    skip_fraction = 0
  out, face_i = [], 0
  for i, w in enumerate(weights):
    if skip_fraction > 0:
      step = skip_fraction/w*len(weights)
      if round(face_i) != round(face_i+step):
        continue
      face_i += step
    out.append((i, _custom_round(w*fan_mode_ray_count)))
  return out


def generate_fan_rays(obj, emit, dist_tol):
  '''
  SurfaceSourceProxy._generateRays(mode='fans') (surface_source.py:469-517) + _makeRay with theta = phi = 0 (:85-111):
  one ray per grid point, leaving along the face normal.  Returns a point_source.RayBatch.
  '''
  from .point_source import RayBatch
  dist_tol = max(float(dist_tol), 1e-9)                                   # _getDistTol, :114-119
  weights = emit.areas/np.sum(emit.areas)
  origins, directions, face_index = [], [], []
  for i, n_rays in fan_ray_counts(weights, float(obj.get('FanModeRayCount', 100))):
    face = FaceEvaluator(emit.faces[i], emit.segs)
    for (u, v), point, (du, dv) in make_surface_grid(face, n_rays, dist_tol):
      normal = face.normal_at(u, v)
      origins.append(point)
      directions.append((point+normal)-point)                             # gpM*pMi*(origin+direction) - origin, :104-106
      face_index.append(i)
  n = len(origins)
  return RayBatch(np.array(origins).reshape(-1, 3), np.array(directions).reshape(-1, 3), np.ones(n), float(obj.get('Wavelength', 500.0)),
                  dict(initPhi=np.zeros(n), initTheta=np.zeros(n), emitFace=np.array(face_index, dtype=np.int64)))
