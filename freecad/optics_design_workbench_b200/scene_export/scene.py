'''
Flat scene description shared by the CUDA engine (through the C ABI in include/odw.h), the
CPU oracle and the tests.  numpy structured dtypes below mirror the C structs byte for byte.

What the reference does per segment — walk every optical group, every placement of it, every
shell, every face, and ask OCC (reference freecad_elements/ray.py:328-432) — is replaced by a
one-shot export: every face instance is written once, in WORLD coordinates, as a closed-form
surface (plane / cylinder / cone / sphere / torus / conic of revolution) plus its trimming region in (u, v) space.
'''

import numpy as np

SURF_PLANE, SURF_CYLINDER, SURF_CONE, SURF_SPHERE, SURF_TORUS, SURF_CONICOID = 1, 2, 3, 4, 5, 6
TRIM_NONE, TRIM_UVBOX, TRIM_LOOPS = 0, 1, 2
SEG_LINE, SEG_ARC, SEG_ASPHERE = 1, 2, 3
OPT_MIRROR, OPT_LENS, OPT_GRATING, OPT_ABSORBER, OPT_VACUUM = 0, 1, 2, 3, 4
OPTICAL_TYPES = ('Mirror', 'Lens', 'Grating', 'Absorber', 'Vacuum')
GRATING_TYPES = ('Reflection', 'Transmission')
SRC_POINT_SPHERICAL, SRC_POINT_COLLIMATED, SRC_SURFACE = 0, 1, 2

_KIND_ID = dict(plane=SURF_PLANE, cylinder=SURF_CYLINDER, cone=SURF_CONE, sphere=SURF_SPHERE,
                torus=SURF_TORUS, conicoid=SURF_CONICOID)

FACE_DTYPE = np.dtype([
  ('origin', '<f8', 3), ('xdir', '<f8', 3), ('ydir', '<f8', 3), ('zdir', '<f8', 3),
  ('p0', '<f8'), ('p1', '<f8'),
  ('uv_min', '<f8', 2), ('uv_max', '<f8', 2),
  ('aabb_min', '<f8', 3), ('aabb_max', '<f8', 3),
  ('kind', '<i4'), ('trim_kind', '<i4'), ('nsign', '<i4'), ('group', '<i4'),
  ('shell', '<i4'), ('seg_first', '<i4'), ('seg_count', '<i4'), ('face_id', '<i4'),
], align=True)
SEG_DTYPE = np.dtype([('a', '<f8', 5), ('kind', '<i4'), ('pad', '<i4')], align=True)
SHELL_DTYPE = np.dtype([
  ('aabb_min', '<f8', 3), ('aabb_max', '<f8', 3),
  ('face_first', '<i4'), ('face_count', '<i4'), ('group', '<i4'), ('pad', '<i4'),
], align=True)
GROUP_DTYPE = np.dtype([
  ('refractive_index', '<f8'), ('reflectivity', '<f8'), ('absorption_length', '<f8'),
  ('grating_lines_per_mm', '<f8'), ('grating_order', '<f8'), ('grating_orientation', '<f8', 3),
  ('optical_type', '<i4'), ('record_hits', '<i4'), ('grating_type', '<i4'), ('fresnel', '<i4'),
], align=True)

assert FACE_DTYPE.itemsize == 224 and SEG_DTYPE.itemsize == 48
assert SHELL_DTYPE.itemsize == 64 and GROUP_DTYPE.itemsize == 80

TWO_PI = 2*np.pi


class UnsupportedGeometry(ValueError):
  'face cannot be expressed in closed form (needs the tessellation path)'


# ------------------------------------------------------------------------------------------
# pcurves -> trim segments

def _curve_to_segs(curve, first, last, deflection=1e-7, out=None):
  out = [] if out is None else out
  kind = curve.kind
  if kind == 'trimmed':
    return _curve_to_segs(curve.basis, first, last, deflection, out)
  if kind == 'line':
    p0 = curve.p + first*curve.d
    p1 = curve.p + last*curve.d
    out.append((SEG_LINE, [p0[0], p0[1], p1[0], p1[1], 0.0]))
    return out
  if kind == 'circle':
    alpha = np.arctan2(curve.dx[1], curve.dx[0])
    sense = np.sign(curve.dx[0]*curve.dy[1]-curve.dx[1]*curve.dy[0])
    span = last-first
    a0 = alpha+first if sense > 0 else alpha-last
    a0 = a0 % TWO_PI
    span = min(span, TWO_PI)
    out.append((SEG_ARC, [curve.p[0], curve.p[1], curve.r, a0, span]))
    return out
  # everything else: polyline fine enough that the chord error is below `deflection`
  n = 32
  while True:
    t = np.linspace(first, last, n+1)
    pts = curve.eval(t)
    tm = (t[1:]+t[:-1])/2
    mid = curve.eval(tm)
    err = np.linalg.norm(mid-(pts[1:]+pts[:-1])/2, axis=-1).max()
    if err < deflection or n >= 4096:
      break
    n *= 2
  for a, b in zip(pts[:-1], pts[1:]):
    out.append((SEG_LINE, [a[0], a[1], b[0], b[1], 0.0]))
  return out


def _segs_bbox(segs):
  lo = np.array([np.inf, np.inf])
  hi = -lo
  for kind, a in segs:
    if kind == SEG_LINE:
      pts = np.array([[a[0], a[1]], [a[2], a[3]]])
    else:
      cu, cv, r, a0, span = a
      angs = [a0, a0+span]
      for k in range(-1, 6):
        ang = k*np.pi/2
        if a0 <= ang <= a0+span:
          angs.append(ang)
      angs = np.array(angs)
      pts = np.stack([cu+r*np.cos(angs), cv+r*np.sin(angs)], axis=-1)
    lo = np.minimum(lo, pts.min(axis=0))
    hi = np.maximum(hi, pts.max(axis=0))
  return lo, hi


def _is_uvbox(segs, lo, hi, eps=1e-9):
  'all segments are axis-aligned lines on the bounding rectangle and together cover its perimeter once'
  total = 0.0
  for kind, a in segs:
    if kind != SEG_LINE:
      return False
    u0, v0, u1, v1 = a[:4]
    if abs(u0-u1) < eps:
      if not (abs(u0-lo[0]) < eps or abs(u0-hi[0]) < eps):
        return False
      total += abs(v1-v0)
    elif abs(v0-v1) < eps:
      if not (abs(v0-lo[1]) < eps or abs(v0-hi[1]) < eps):
        return False
      total += abs(u1-u0)
    else:
      return False
  return abs(total-2*((hi[0]-lo[0])+(hi[1]-lo[1]))) < 1e-6*max(1.0, total)


def point_in_segs(segs, u, v):
  '''
  Even-odd test of (u, v) against trim segments given as (kind, a) tuples or SEG_DTYPE rows:
  count crossings of the half-line {(u', v): u' > u}.  numpy restatement used by tests/export;
  the CUDA kernel (csrc/odw_device.cuh loops_contain) implements the same rule.
  '''
  crossings = 0
  for s in segs:
    kind, a = (s['kind'], s['a']) if isinstance(s, np.void) else s
    if kind == SEG_LINE:
      u0, v0, u1, v1 = a[:4]
      if (v0 > v) != (v1 > v):
        ux = u0 + (v-v0)*(u1-u0)/(v1-v0)
        if ux > u:
          crossings += 1
    else:
      cu, cv, r, a0, span = a[:5]
      dv = v-cv
      if abs(dv) < r:
        h = np.sqrt(r*r-dv*dv)
        for ux in (cu-h, cu+h):
          if ux > u:
            ang = np.arctan2(dv, ux-cu)
            rel = (ang-a0) % TWO_PI
            if rel <= span:
              crossings += 1
  return (crossings & 1) == 1


# ------------------------------------------------------------------------------------------
# face instance -> FACE_DTYPE row

def _rigid(transform):
  R = transform[:3, :3]
  s = np.cbrt(abs(np.linalg.det(R)))
  if abs(s-1) > 1e-9:
    raise UnsupportedGeometry(f'non-rigid placement (scale {s}) is not supported')
  return R, transform[:3, 3]


def conic_sag(c, k, rho, poly=None):
  '''
  optical sag of a conic of revolution, z = c rho^2 / (1 + sqrt(1 - (1+k) c^2 rho^2)), plus the even-asphere terms
  poly[j] rho^(2j+4), j = 0..4 (include/odw.h ODW_SEG_ASPHERE)
  '''
  rho = np.asarray(rho, dtype=float)
  u = rho*rho
  z = c*u/(1+np.sqrt(np.maximum(0.0, 1-(1+k)*c*c*u)))
  if poly is not None:
    a = list(poly)+[0.0]*(5-len(poly))
    z = z + u*u*(a[0] + u*(a[1] + u*(a[2] + u*(a[3] + u*a[4]))))
  return z


def _revolved_conic(surf):
  '''
  A surface of revolution that is a conic of revolution in closed form, as a 'conicoid' Surface (frame + c, k), or None.
  Recognised: a Geom_Parabola revolved about its own axis (what Part.Parabola + Revolve writes) = paraboloid, k = -1,
  c = 1/(2 F).  OCC's (u, v) = (rotation angle about the axis, parabola parameter t = signed distance from the axis):
  the frame is chosen so that v = t = rho on the t >= 0 branch.
  '''
  from .brep import Surface
  curve = surf.curve
  while curve.kind == 'trimmed':
    curve = curve.basis
  if curve.kind != 'parabola':
    return None
  axis = np.asarray(surf.d, dtype=float)
  axis = axis/np.linalg.norm(axis)
  Z = np.asarray(curve.dx, dtype=float)                  # symmetry axis of the parabola, apex -> focus
  if np.linalg.norm(np.cross(axis, Z)) > 1e-12:
    return None                                           # revolved about another line: not a paraboloid
  off = np.asarray(curve.p, dtype=float)-np.asarray(surf.p, dtype=float)
  if np.linalg.norm(off-np.dot(off, axis)*axis) > 1e-9*max(1.0, abs(curve.f)):
    return None                                           # the axis of revolution misses the apex
  X = np.asarray(curve.dy, dtype=float)
  Y = np.cross(axis, X)                                   # rotation by u about `axis` takes X towards Y
  return Surface('conicoid', p=np.asarray(curve.p, dtype=float), n=Z, dx=X, dy=Y, c=1.0/(2.0*curve.f), k=-1.0)


ASPHERE_FIT_TOLERANCE = 1e-7      # mm: largest sag residual for which a revolved free-form meridian is taken as an even asphere


def _revolved_asphere(surf, tolerance=None):
  '''
  A surface of revolution whose meridian is a free-form curve (B-spline / Bezier — how FreeCAD users model an aspheric lens:
  a spline through points of the lens formula, revolved about the optical axis) fitted to the even-asphere form
    z = z0 + c rho^2 / (1 + sqrt(1 - (1+k) c^2 rho^2)) + a4 rho^4 + ... + a12 rho^12        (include/odw.h ODW_SEG_ASPHERE)
  Returns a 'conicoid' Surface with .poly and .vmap (curve parameter t -> rho, for the trim curves), or None when the
  meridian is not of that kind or the fit leaves a sag residual above `tolerance` (then the face is meshed as before, with
  the tessellation's deflection as its error).  The residual is reported in .fit_residual.
  '''
  from .brep import Surface
  import scipy.optimize
  tolerance = ASPHERE_FIT_TOLERANCE if tolerance is None else tolerance
  curve, t0, t1 = surf.curve, None, None
  while curve.kind == 'trimmed':
    t0 = curve.u1 if t0 is None else t0
    t1 = curve.u2 if t1 is None else t1
    curve = curve.basis
  if curve.kind not in ('bspline', 'bezier'):
    return None
  spline = curve.spline
  if t0 is None:
    t0, t1 = float(spline.knots[0]), float(spline.knots[-1])
  axis = np.asarray(surf.d, dtype=float)
  axis = axis/np.linalg.norm(axis)
  A = np.asarray(surf.p, dtype=float)
  t = np.linspace(t0, t1, 801)
  w = curve.eval(t)-A
  z = w @ axis
  radial = w-z[:, None]*axis
  rho = np.linalg.norm(radial, axis=1)
  far = int(np.argmax(rho))
  if rho[far] < 1e-9:
    return None
  X = radial[far]/rho[far]
  if np.abs(radial-rho[:, None]*X).max() > 1e-9*max(1.0, rho[far]):
    return None                                           # the meridian leaves the half plane through the axis (or crosses the axis)
  if np.any(np.diff(rho) <= 0) and np.any(np.diff(rho) >= 0):
    return None                                           # rho must be monotonic along the curve: one sag value per radius
  order = np.argsort(rho)
  r, zz = rho[order], z[order]

  def sag(params, r):
    z0, c, k = params[:3]
    u = r*r
    root = np.sqrt(np.maximum(1e-300, 1-(1+k)*c*c*u))
    a = params[3:]
    return z0 + c*u/(1+root) + u*u*(a[0] + u*(a[1] + u*(a[2] + u*(a[3] + u*a[4]))))

  # start: vertex height by extrapolation, vertex curvature from the innermost part of the meridian
  inner = r <= max(r[0] + 0.2*(r[-1]-r[0]), r[min(len(r)-1, 8)])
  quad = np.polyfit(r[inner]**2, zz[inner], 1)
  start = np.array([quad[1], 2*quad[0] if quad[0] != 0 else 1e-6, 0.0, 0, 0, 0, 0, 0], dtype=float)
  scale = np.array([1.0, 1.0, 1.0] + [r[-1]**-(2*j+4) for j in range(5)])       # polynomial terms of order one at the rim
  best = None
  for n_poly in (0, 2, 5):                                # plain conic first, then with rho^4, rho^6, then all five terms
    def residual(q, n_poly=n_poly):
      params = np.concatenate([q[:3], q[3:3+n_poly]*scale[3:3+n_poly], np.zeros(5-n_poly)])
      if (1+params[2])*params[1]**2*r[-1]**2 >= 1:
        return np.full_like(r, 1e3)
      return sag(params, r)-zz
    q0 = np.concatenate([start[:3], np.zeros(n_poly)]) if best is None else np.concatenate([best[0][:3], (best[0][3:]/scale[3:])[:n_poly]])
    sol = scipy.optimize.least_squares(residual, q0, xtol=1e-15, ftol=1e-15, gtol=1e-15, max_nfev=2000)
    params = np.concatenate([sol.x[:3], sol.x[3:3+n_poly]*scale[3:3+n_poly], np.zeros(5-n_poly)])
    err = float(np.abs(sag(params, r)-zz).max())
    if best is None or err < best[1]:
      best = (params, err)
    if err <= tolerance:
      break
  params, err = best
  if not (err <= tolerance) or not np.isfinite(params).all() or params[1] == 0:
    return None
  z0, c, k = params[:3]
  Y = np.cross(axis, X)                                   # OCC: rotation by u about the axis takes the meridian plane from X towards Y
  out = Surface('conicoid', p=A+z0*axis, n=axis, dx=X, dy=Y, c=float(c), k=float(k),
                poly=[float(a) for a in params[3:]] if np.any(params[3:] != 0) else None)
  out.fit_residual = err
  tt, rr = t, rho
  if rr[0] > rr[-1]:
    tt, rr = tt[::-1], rr[::-1]
  out.vmap = lambda v: float(np.linalg.norm((lambda q: q-(q @ axis)*axis)(curve.eval(np.array([v]))[0]-A)))
  return out


def _map_v(segs, vmap):
  'trim pieces given in (u, t) of a revolved curve -> (u, rho): pointwise on straight pieces (split where both coordinates change)'
  out = []
  for kind, a in segs:
    if kind != SEG_LINE:
      raise UnsupportedGeometry('curved trim piece on a fitted surface of revolution')
    u0, v0, u1, v1 = a[:4]
    n = 1 if (abs(u1-u0) < 1e-12 or abs(v1-v0) < 1e-12) else 32
    for j in range(n):
      ua, ub = u0+(u1-u0)*j/n, u0+(u1-u0)*(j+1)/n
      va, vb = v0+(v1-v0)*j/n, v0+(v1-v0)*(j+1)/n
      out.append((SEG_LINE, [ua, vmap(va), ub, vmap(vb), 0.0]))
  return out


def face_record(fi, transform, group, shell, face_id, segs_out):
  '''
  Convert a brep.FaceInstance (+ an extra world transform applied on the left) into a FACE_DTYPE
  row; trim segments are appended to segs_out (list of (kind, a)).
  '''
  surf = fi.surface
  if surf.kind == 'revolution':
    surf = _revolved_conic(surf) or _revolved_asphere(surf) or surf
  if surf.kind not in _KIND_ID:
    raise UnsupportedGeometry(f'surface kind {surf.kind!r} has no closed form')
  R, T = _rigid(transform @ fi.transform)
  f = np.zeros((), dtype=FACE_DTYPE)
  f['origin'] = R @ surf.p + T
  X, Y, Z = R @ surf.dx, R @ surf.dy, R @ surf.n
  f['xdir'], f['ydir'], f['zdir'] = X, Y, Z
  handed = 1 if np.dot(np.cross(X, Y), Z) > 0 else -1
  f['nsign'] = (-1 if fi.reversed else 1)*handed
  f['kind'] = _KIND_ID[surf.kind]
  if surf.kind in ('cylinder', 'sphere'):
    f['p0'] = surf.r
  elif surf.kind == 'cone':
    f['p0'], f['p1'] = surf.r, surf.angle
  elif surf.kind == 'torus':
    f['p0'], f['p1'] = surf.r, surf.r2
  elif surf.kind == 'conicoid':
    if not (np.isfinite(surf.c) and surf.c != 0 and np.isfinite(surf.k)):
      raise UnsupportedGeometry('conicoid needs a finite non-zero vertex curvature and a finite conic constant')
    f['p0'], f['p1'] = surf.c, surf.k
  f['group'], f['shell'], f['face_id'] = group, shell, face_id
  poly = [float(x) for x in (getattr(surf, 'poly', None) or [])] if surf.kind == 'conicoid' else []
  if len(poly) > 5 or not all(np.isfinite(poly)):
    raise ValueError('a conicoid takes up to five finite even-asphere coefficients (rho^4 .. rho^12)')
  if not any(poly):
    poly = []

  segs = []
  for loop in fi.loops:
    for curve, first, last in loop:
      _curve_to_segs(curve, first, last, out=segs)
  if getattr(surf, 'vmap', None) is not None:
    segs = _map_v(segs, surf.vmap)                        # the face's v is the parameter of the revolved curve, the conicoid's is rho
  if not segs:
    if surf.kind in ('sphere', 'torus'):
      lo, hi = np.array([0.0, -np.pi/2 if surf.kind == 'sphere' else 0.0]), \
               np.array([TWO_PI, np.pi/2 if surf.kind == 'sphere' else TWO_PI])
      trim = TRIM_NONE
    else:
      raise UnsupportedGeometry('face without boundary on an unbounded surface')
  else:
    lo, hi = _segs_bbox(segs)
    trim = TRIM_UVBOX if _is_uvbox(segs, lo, hi) else TRIM_LOOPS
    if surf.kind == 'conicoid':
      if lo[1] < -1e-9*max(1.0, abs(hi[1])):
        raise UnsupportedGeometry('conicoid face on the negative branch of its meridian')
      lo[1] = max(lo[1], 0.0)
      if (1+surf.k)*surf.c**2*hi[1]**2 > 1:
        raise ValueError('conicoid face reaches beyond the equator of its ellipsoid: the sag formula has no value there')
    if trim == TRIM_UVBOX:
      full_u = abs((hi[0]-lo[0])-TWO_PI) < 1e-9
      if surf.kind == 'sphere' and full_u and lo[1] < -np.pi/2+1e-9 and hi[1] > np.pi/2-1e-9:
        trim = TRIM_NONE
      if surf.kind == 'torus' and full_u and abs((hi[1]-lo[1])-TWO_PI) < 1e-9:
        trim = TRIM_NONE
  f['trim_kind'] = trim
  f['uv_min'], f['uv_max'] = lo, hi
  aux = [(SEG_ASPHERE, poly+[0.0]*(5-len(poly)))] if poly else []       # auxiliary record first, boundary pieces after it
  if trim == TRIM_LOOPS or aux:
    f['seg_first'], f['seg_count'] = len(segs_out), len(aux)+(len(segs) if trim == TRIM_LOOPS else 0)
    segs_out.extend(aux)
    if trim == TRIM_LOOPS:
      segs_out.extend(segs)
  f['aabb_min'], f['aabb_max'] = _face_aabb(f, poly=poly or None)
  return f


def eval_face(f, u, v, poly=None):
  'world point(s) of FACE_DTYPE row f at parameters u, v (poly: even-asphere coefficients of a conicoid face)'
  u = np.asarray(u, dtype=float)[..., None]
  v = np.asarray(v, dtype=float)[..., None]
  O, X, Y, Z = f['origin'], f['xdir'], f['ydir'], f['zdir']
  k = int(f['kind'])
  if k == SURF_PLANE:
    return O + u*X + v*Y
  er = np.cos(u)*X + np.sin(u)*Y
  if k == SURF_CYLINDER:
    return O + f['p0']*er + v*Z
  if k == SURF_CONE:
    return O + (f['p0']+v*np.sin(f['p1']))*er + v*np.cos(f['p1'])*Z
  if k == SURF_SPHERE:
    return O + f['p0']*np.cos(v)*er + f['p0']*np.sin(v)*Z
  if k == SURF_TORUS:
    return O + (f['p0']+f['p1']*np.cos(v))*er + f['p1']*np.sin(v)*Z
  if k == SURF_CONICOID:
    return O + v*er + conic_sag(float(f['p0']), float(f['p1']), v, poly)*Z
  raise ValueError(k)


def _face_aabb(f, n=65, poly=None):
  lo, hi = f['uv_min'], f['uv_max']
  k = int(f['kind'])
  if k == SURF_PLANE:
    uu, vv = np.meshgrid([lo[0], hi[0]], [lo[1], hi[1]])
    pts = eval_face(f, uu, vv).reshape(-1, 3)
    return pts.min(axis=0), pts.max(axis=0)
  uu, vv = np.meshgrid(np.linspace(lo[0], hi[0], n), np.linspace(lo[1], hi[1], n))
  pts = eval_face(f, uu, vv, poly).reshape(-1, 3)
  # sagitta of the sampling: a point between samples can stick out by R(1-cos(h/2))
  hu = (hi[0]-lo[0])/(n-1)
  rmax = {SURF_CYLINDER: f['p0'], SURF_SPHERE: f['p0'], SURF_TORUS: f['p0']+f['p1'],
          SURF_CONE: abs(f['p0'])+max(abs(lo[1]), abs(hi[1]))*abs(np.sin(f['p1'])),
          SURF_CONICOID: hi[1]}[k]
  pad = rmax*(1-np.cos(hu/2))
  if k == SURF_CONICOID:
    # between two samples of a meridian the sag deviates from the chord by at most h^2/8 max|z''|, z'' = c / q^3
    hv = (hi[1]-lo[1])/(n-1)
    q = np.sqrt(max(1e-12, 1-(1+float(f['p1']))*float(f['p0'])**2*hi[1]**2))
    pad += hv*hv/8*abs(float(f['p0']))/q**3
    if poly is not None:                     # second derivative of the polynomial part, bounded on a fine grid
      r = np.linspace(lo[1], hi[1], 1025)
      z = conic_sag(0.0, 0.0, r, poly)
      pad += hv*hv/8*1.5*np.abs(np.gradient(np.gradient(z, r), r)).max()
  if k in (SURF_SPHERE, SURF_TORUS):
    hv = (hi[1]-lo[1])/(n-1)
    rv = f['p0'] if k == SURF_SPHERE else f['p1']
    pad += rv*(1-np.cos(hv/2))
  return pts.min(axis=0)-pad, pts.max(axis=0)+pad


# ------------------------------------------------------------------------------------------

class Scene:
  '''
  Immutable scene: numpy arrays laid out exactly like the C structs of include/odw.h plus the
  names needed by the hit writer.
  '''
  def __init__(self, faces, segs, shells, groups, group_names, group_labels,
               seq_offsets=None, seq_groups=None, scatters=None, group_scatter=None):
    self.faces = np.ascontiguousarray(faces, dtype=FACE_DTYPE)
    self.segs = np.ascontiguousarray(segs, dtype=SEG_DTYPE)
    self.shells = np.ascontiguousarray(shells, dtype=SHELL_DTYPE)
    self.groups = np.ascontiguousarray(groups, dtype=GROUP_DTYPE)
    self.group_names = list(group_names)
    self.group_labels = list(group_labels)
    self.seq_offsets = np.ascontiguousarray(seq_offsets if seq_offsets is not None else [0], dtype=np.int32)
    self.seq_groups = np.ascontiguousarray(seq_groups if seq_groups is not None else [], dtype=np.int32)
    # stochastic surface models: list of distributions.SamplerTables + per group {main, modify} index or -1
    self.scatters = list(scatters or [])
    self.group_scatter = (np.ascontiguousarray(group_scatter, dtype=np.int32).reshape(len(self.groups), 2)
                          if group_scatter is not None else np.full((len(self.groups), 2), -1, dtype=np.int32))

  @property
  def n_seq_steps(self):
    return len(self.seq_offsets)-1

  def summary(self):
    kinds = {v: k for k, v in _KIND_ID.items()}
    from collections import Counter
    c = Counter((kinds[int(f['kind'])], ('none', 'uvbox', 'loops')[int(f['trim_kind'])]) for f in self.faces)
    return dict(faces=len(self.faces), segs=len(self.segs), shells=len(self.shells),
                groups=len(self.groups), census={f'{k}/{t}': n for (k, t), n in sorted(c.items())})


class SceneBuilder:
  'collects face instances group by group, shell by shell'

  def __init__(self):
    self.faces, self.segs, self.shells, self.groups = [], [], [], []
    self.group_names, self.group_labels = [], []
    self.scatters, self.group_scatter = [], []
    self.skipped = []          # (group, reason) for faces that could neither be written in closed form nor meshed
    self.tessellated = []      # one entry per face that went through the tessellation path (grid, achieved deflection)
    from .tessellate import DEFAULT_DEFLECTION
    self.deflection = DEFAULT_DEFLECTION

  def add_group(self, name, label, optical_type, refractive_index=1.0, reflectivity=1.0,
                absorption_length=np.inf, record_hits=False, grating_type=0,
                grating_lines_per_mm=1000.0, grating_order=1.0, grating_orientation=(0, 0, 1),
                scatter_density='', power_theta_domain='-pi/2, pi/2', power_phi_domain='0, 2*pi',
                modify_density='', modify_theta_domain='-pi/2, pi/2', modify_phi_domain='0, 2*pi',
                scatter_resolution=None, fresnel=False):
    '''
    fresnel = opt-in Fresnel reflection at the faces of a Lens group (include/odw.h odw_group.fresnel; the reference has
    none, so the default reproduces it).
    scatter_density = ReflectedProbabilityDensity (Mirror) / RefractedProbabilityDensity (Lens); modify_density =
    RayModificationProbabilityDensity (optical_group.py:29-96); empty = ideal surface.
    '''
    g = np.zeros((), dtype=GROUP_DTYPE)
    g['optical_type'] = (OPTICAL_TYPES.index(optical_type) if isinstance(optical_type, str)
                         else optical_type)
    g['refractive_index'], g['reflectivity'] = refractive_index, reflectivity
    g['absorption_length'] = absorption_length
    g['record_hits'] = int(bool(record_hits))
    g['fresnel'] = int(bool(fresnel))
    g['grating_type'] = (GRATING_TYPES.index(grating_type) if isinstance(grating_type, str)
                         else grating_type)
    g['grating_lines_per_mm'], g['grating_order'] = grating_lines_per_mm, grating_order
    g['grating_orientation'] = grating_orientation
    self.groups.append(g)
    self.group_names.append(name)
    self.group_labels.append(label)
    from ..distributions import scatter_tables
    idx = [-1, -1]
    if int(g['optical_type']) in (OPT_MIRROR, OPT_LENS):       # applyStochasticRayCorrections is only called for these
      for k, (dens, td, pd) in enumerate(((scatter_density, power_theta_domain, power_phi_domain),
                                          (modify_density, modify_theta_domain, modify_phi_domain))):
        # only the main density is re-compiled per hit by the reference (optical_group.py:307); the modify density is drawn as it is (:318)
        t = (scatter_tables(dens, td, pd, scatter_resolution, optical_type=OPTICAL_TYPES[int(g['optical_type'])],
                            refractive_index=float(g['refractive_index'])) if k == 0 else scatter_tables(dens, td, pd, scatter_resolution))
        if t is not None:
          self.scatters.append(t)
          idx[k] = len(self.scatters)-1
    self.group_scatter.append(idx)
    return len(self.groups)-1

  def add_shape(self, group, face_instances, transform):
    '''
    face_instances: brep.FaceInstance list of ONE shape; transform: 4x4 world matrix applied on top.
    Faces are grouped into shells by FaceInstance.shell_key; like the reference
    (`cachedShells(shape) or [shape]`, ray.py:345) faces outside any shell only count when the shape has
    no shell at all.
    '''
    keys = []
    for fi in face_instances:
      if fi.shell_key not in keys:
        keys.append(fi.shell_key)
    if any(k is not None for k in keys):
      keys = [k for k in keys if k is not None]
    for key in keys:
      shell_index = len(self.shells)
      first = len(self.faces)
      for fi in face_instances:
        if fi.shell_key != key:
          continue
        try:
          self.faces.append(face_record(fi, transform, group, shell_index, len(self.faces), self.segs))
        except UnsupportedGeometry as e:
          # no closed form: tessellate (scene_export/tessellate.py); only what cannot be meshed either is skipped
          try:
            from . import tessellate
            tris, info = tessellate.triangle_faces(fi, transform, group, shell_index, len(self.faces), self.segs,
                                                   deflection=self.deflection)
            self.faces.extend(tris)
            self.tessellated.append(dict(group=group, reason=str(e), **info))
          except UnsupportedGeometry as e2:
            self.skipped.append((group, f'{e}; {e2}'))
      count = len(self.faces)-first
      if count == 0:
        continue
      sh = np.zeros((), dtype=SHELL_DTYPE)
      fa = self.faces[first:]
      sh['aabb_min'] = np.min([f['aabb_min'] for f in fa], axis=0)
      sh['aabb_max'] = np.max([f['aabb_max'] for f in fa], axis=0)
      sh['face_first'], sh['face_count'], sh['group'] = first, count, group
      self.shells.append(sh)

  def build(self, sequence=None):
    segs = np.zeros(len(self.segs), dtype=SEG_DTYPE)
    for i, (kind, a) in enumerate(self.segs):
      segs[i]['kind'] = kind
      segs[i]['a'] = a
    faces = np.array(self.faces, dtype=FACE_DTYPE) if self.faces else np.zeros(0, dtype=FACE_DTYPE)
    shells = np.array(self.shells, dtype=SHELL_DTYPE) if self.shells else np.zeros(0, dtype=SHELL_DTYPE)
    groups = np.array(self.groups, dtype=GROUP_DTYPE) if self.groups else np.zeros(0, dtype=GROUP_DTYPE)
    seq_offsets, seq_groups = [0], []
    for step in (sequence or []):
      seq_groups.extend(step)
      seq_offsets.append(len(seq_groups))
    return Scene(faces, segs, shells, groups, self.group_names, self.group_labels, seq_offsets, seq_groups,
                 scatters=self.scatters, group_scatter=self.group_scatter if self.group_scatter else None)
