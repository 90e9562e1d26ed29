'''
Reader for OpenCASCADE ASCII BRep files (`*.Shape.brp` members of a FCStd archive).

The reference never parses BRep itself: it asks FreeCAD/OCC for `Shape.Shells`, `shell.Faces`,
`face.Surface`, `face.normalAt`, `vertex.distToShape(face)` (reference freecad_elements/ray.py:345-426,
455-480).  The scene export needs the same information without FreeCAD, so this module reads what
OCC wrote: the location table, the 2-D parametric curves (pcurves), the surfaces and the topology
(TShapes).  Output: for every face instance its surface, rigid transform, orientation and the
boundary loops expressed in the surface's own (u, v) space.

Format notes (BRepTools_ShapeSet / GeomTools text format, "CASCADE Topology V1/V2/V3"):
sub-shape index k of a file with N TShapes refers to the (N-k+1)-th TShape in file order.
'''

import re
import numpy as np

_SECTION_KEYS = ('Locations', 'Curve2ds', 'Curves', 'Polygon3D', 'PolygonOnTriangulations',
                 'Surfaces', 'Triangulations', 'TShapes')
_CONT = ('C0', 'C1', 'C2', 'C3', 'CN', 'G1', 'G2')


class BRepError(ValueError):
  pass


class _Tokens:
  def __init__(self, toks, pos=0, end=None):
    self.t = toks
    self.i = pos
    self.end = len(toks) if end is None else end

  def more(self):
    return self.i < self.end

  def peek(self):
    return self.t[self.i]

  def next(self):
    v = self.t[self.i]
    self.i += 1
    return v

  def int(self):
    return int(self.next())

  def real(self):
    return float(self.next())

  def reals(self, n):
    v = [float(x) for x in self.t[self.i:self.i+n]]
    self.i += n
    return np.array(v)


# ------------------------------------------------------------------------------------------
# 2-D curves (pcurves)

class Line2d:
  kind = 'line'
  def __init__(self, p, d):
    self.p, self.d = p, d
  def eval(self, t):
    t = np.asarray(t, dtype=float)
    return self.p + t[..., None]*self.d


class Circle2d:
  kind = 'circle'
  def __init__(self, p, dx, dy, r):
    self.p, self.dx, self.dy, self.r = p, dx, dy, r
  def eval(self, t):
    t = np.asarray(t, dtype=float)
    return self.p + self.r*(np.cos(t)[..., None]*self.dx + np.sin(t)[..., None]*self.dy)


class Ellipse2d:
  kind = 'ellipse'
  def __init__(self, p, dx, dy, r1, r2):
    self.p, self.dx, self.dy, self.r1, self.r2 = p, dx, dy, r1, r2
  def eval(self, t):
    t = np.asarray(t, dtype=float)
    return self.p + self.r1*np.cos(t)[..., None]*self.dx + self.r2*np.sin(t)[..., None]*self.dy


class Parabola2d:
  kind = 'parabola'
  def __init__(self, p, dx, dy, f):
    self.p, self.dx, self.dy, self.f = p, dx, dy, f
  def eval(self, t):
    t = np.asarray(t, dtype=float)
    return self.p + (t*t/(4*self.f))[..., None]*self.dx + t[..., None]*self.dy


class Hyperbola2d:
  kind = 'hyperbola'
  def __init__(self, p, dx, dy, r1, r2):
    self.p, self.dx, self.dy, self.r1, self.r2 = p, dx, dy, r1, r2
  def eval(self, t):
    t = np.asarray(t, dtype=float)
    return self.p + self.r1*np.cosh(t)[..., None]*self.dx + self.r2*np.sinh(t)[..., None]*self.dy


def _find_span(knots, degree, npoles, u):
  # knots: flat (expanded) knot vector
  lo, hi = degree, npoles
  if u >= knots[hi]:
    return hi-1
  if u <= knots[lo]:
    return lo
  return int(np.searchsorted(knots, u, side='right')-1)


def _de_boor(flat_knots, degree, ctrl, u):
  'evaluate a (homogeneous) B-spline with control points ctrl[n,dim] at scalar u'
  n = len(ctrl)
  k = _find_span(flat_knots, degree, n, u)
  d = [np.array(ctrl[j+k-degree], dtype=float) for j in range(degree+1)]
  for r in range(1, degree+1):
    for j in range(degree, r-1, -1):
      i = j+k-degree
      den = flat_knots[i+degree-r+1]-flat_knots[i]
      a = 0.0 if den == 0 else (u-flat_knots[i])/den
      d[j] = (1-a)*d[j-1] + a*d[j]
  return d[degree]


class BSplineCurve:
  'rational/non-rational, periodic or not, any dimension'
  kind = 'bspline'
  def __init__(self, degree, poles, weights, knots, mults, periodic):
    self.degree = degree
    self.poles = np.asarray(poles, dtype=float)
    self.weights = None if weights is None else np.asarray(weights, dtype=float)
    self.knots = np.asarray(knots, dtype=float)
    self.mults = np.asarray(mults, dtype=int)
    self.periodic = periodic
    poles_h = (self.poles if self.weights is None
               else np.hstack([self.poles*self.weights[:, None], self.weights[:, None]]))
    self.flat, self.period, idx = _flat_knots(self.knots, self.mults, degree, periodic, len(poles_h))
    self.ctrl = poles_h[idx]

  def eval(self, t):
    t = np.atleast_1d(np.asarray(t, dtype=float))
    out = []
    for u in t.ravel():
      if self.period is not None:
        u = self.knots[0] + (u-self.knots[0]) % self.period
      p = _de_boor(self.flat, self.degree, self.ctrl, u)
      if self.weights is not None:
        p = p[:-1]/p[-1]
      out.append(p)
    return np.array(out).reshape(t.shape+(self.poles.shape[1],))


class BezierCurve(BSplineCurve):
  kind = 'bezier'
  def __init__(self, poles, weights):
    deg = len(poles)-1
    super().__init__(deg, poles, weights, [0.0, 1.0], [deg+1, deg+1], False)


class TrimmedCurve:
  kind = 'trimmed'
  def __init__(self, u1, u2, basis):
    self.u1, self.u2, self.basis = u1, u2, basis
  def eval(self, t):
    return self.basis.eval(t)


class OffsetCurve2d:
  kind = 'offset'
  def __init__(self, offset, basis):
    self.offset, self.basis = offset, basis
  def eval(self, t):
    t = np.asarray(t, dtype=float)
    h = 1e-6
    p = self.basis.eval(t)
    d = (self.basis.eval(t+h)-self.basis.eval(t-h))/(2*h)
    n = np.stack([d[..., 1], -d[..., 0]], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    return p + self.offset*n


def _read_poles(tk, n, dim, rational):
  poles, weights = [], ([] if rational else None)
  for _ in range(n):
    poles.append(tk.reals(dim))
    if rational:
      weights.append(tk.real())
  return np.array(poles), (None if weights is None else np.array(weights))


def _read_bspline_curve(tk, dim):
  rational, periodic = tk.int(), tk.int()
  degree, npoles, nknots = tk.int(), tk.int(), tk.int()
  poles, weights = _read_poles(tk, npoles, dim, rational)
  knots, mults = [], []
  for _ in range(nknots):
    knots.append(tk.real())
    mults.append(tk.int())
  return BSplineCurve(degree, poles, weights, knots, mults, bool(periodic))


def _read_curve2d(tk):
  typ = tk.int()
  if typ == 1:
    return Line2d(tk.reals(2), tk.reals(2))
  if typ == 2:
    return Circle2d(tk.reals(2), tk.reals(2), tk.reals(2), tk.real())
  if typ == 3:
    return Ellipse2d(tk.reals(2), tk.reals(2), tk.reals(2), tk.real(), tk.real())
  if typ == 4:
    return Parabola2d(tk.reals(2), tk.reals(2), tk.reals(2), tk.real())
  if typ == 5:
    return Hyperbola2d(tk.reals(2), tk.reals(2), tk.reals(2), tk.real(), tk.real())
  if typ == 6:
    rational, degree = tk.int(), tk.int()
    poles, weights = _read_poles(tk, degree+1, 2, rational)
    return BezierCurve(poles, weights)
  if typ == 7:
    return _read_bspline_curve(tk, 2)
  if typ == 8:
    u1, u2 = tk.real(), tk.real()
    return TrimmedCurve(u1, u2, _read_curve2d(tk))
  if typ == 9:
    off = tk.real()
    return OffsetCurve2d(off, _read_curve2d(tk))
  raise BRepError(f'unknown 2d curve type {typ}')


# ------------------------------------------------------------------------------------------
# 3-D curves (only needed as generatrix of extrusion / revolution surfaces)

class Curve3d:
  def __init__(self, kind, **kw):
    self.kind = kind
    self.__dict__.update(kw)

  def eval(self, t):
    t = np.asarray(t, dtype=float)
    k = self.kind
    if k == 'line':
      return self.p + t[..., None]*self.d
    if k == 'circle':
      return self.p + self.r*(np.cos(t)[..., None]*self.dx + np.sin(t)[..., None]*self.dy)
    if k == 'ellipse':
      return self.p + self.r1*np.cos(t)[..., None]*self.dx + self.r2*np.sin(t)[..., None]*self.dy
    if k == 'parabola':
      return self.p + (t*t/(4*self.f))[..., None]*self.dx + t[..., None]*self.dy
    if k == 'hyperbola':
      return self.p + self.r1*np.cosh(t)[..., None]*self.dx + self.r2*np.sinh(t)[..., None]*self.dy
    if k in ('bspline', 'bezier'):
      return self.spline.eval(t)
    if k == 'trimmed':
      return self.basis.eval(t)
    raise BRepError(f'cannot evaluate 3d curve kind {k}')


def _read_curve3d(tk):
  typ = tk.int()
  if typ == 1:
    return Curve3d('line', p=tk.reals(3), d=tk.reals(3))
  if typ == 2:
    return Curve3d('circle', p=tk.reals(3), n=tk.reals(3), dx=tk.reals(3), dy=tk.reals(3), r=tk.real())
  if typ == 3:
    return Curve3d('ellipse', p=tk.reals(3), n=tk.reals(3), dx=tk.reals(3), dy=tk.reals(3),
                   r1=tk.real(), r2=tk.real())
  if typ == 4:
    return Curve3d('parabola', p=tk.reals(3), n=tk.reals(3), dx=tk.reals(3), dy=tk.reals(3), f=tk.real())
  if typ == 5:
    return Curve3d('hyperbola', p=tk.reals(3), n=tk.reals(3), dx=tk.reals(3), dy=tk.reals(3),
                   r1=tk.real(), r2=tk.real())
  if typ == 6:
    rational, degree = tk.int(), tk.int()
    poles, weights = _read_poles(tk, degree+1, 3, rational)
    return Curve3d('bezier', spline=BezierCurve(poles, weights))
  if typ == 7:
    return Curve3d('bspline', spline=_read_bspline_curve(tk, 3))
  if typ == 8:
    u1, u2 = tk.real(), tk.real()
    return Curve3d('trimmed', u1=u1, u2=u2, basis=_read_curve3d(tk))
  if typ == 9:
    off, d = tk.real(), tk.reals(3)
    return Curve3d('offset', offset=off, d=d, basis=_read_curve3d(tk))
  raise BRepError(f'unknown 3d curve type {typ}')


# ------------------------------------------------------------------------------------------
# surfaces

class Surface:
  '''
  kind in plane/cylinder/cone/sphere/torus carries frame (p, n, dx, dy) and radii;
  other kinds (extrusion, revolution, bezier, bspline, offset) carry what is needed to evaluate.
  '''
  ELEMENTARY = {'plane': 1, 'cylinder': 2, 'cone': 3, 'sphere': 4, 'torus': 5}

  def __init__(self, kind, **kw):
    self.kind = kind
    self.__dict__.update(kw)

  def is_elementary(self):
    return self.kind in Surface.ELEMENTARY

  def eval(self, u, v):
    'point(s) on the surface for parameter arrays u, v (broadcast)'
    u = np.asarray(u, dtype=float)
    v = np.asarray(v, dtype=float)
    k = self.kind
    if k == 'rtrimmed' or k == 'offsetsurf_zero':
      return self.basis.eval(u, v)
    if k in Surface.ELEMENTARY:
      X, Y, Z, O = self.dx, self.dy, self.n, self.p
      cu, su = np.cos(u)[..., None], np.sin(u)[..., None]
      if k == 'plane':
        return O + u[..., None]*X + v[..., None]*Y
      if k == 'cylinder':
        return O + self.r*(cu*X + su*Y) + v[..., None]*Z
      if k == 'cone':
        rho = (self.r + v*np.sin(self.angle))[..., None]
        return O + rho*(cu*X + su*Y) + (v*np.cos(self.angle))[..., None]*Z
      if k == 'sphere':
        cv, sv = np.cos(v)[..., None], np.sin(v)[..., None]
        return O + self.r*cv*(cu*X + su*Y) + self.r*sv*Z
      if k == 'torus':
        cv, sv = np.cos(v)[..., None], np.sin(v)[..., None]
        return O + (self.r + self.r2*cv)*(cu*X + su*Y) + self.r2*sv*Z
    if k == 'extrusion':
      c = self.curve.eval(u)
      return c + v[..., None]*self.d
    if k == 'revolution':
      # rotate generatrix point C(v) by angle u about axis (p, d)
      uu, vv = np.broadcast_arrays(u, v)
      c = self.curve.eval(vv)
      a = self.d/np.linalg.norm(self.d)
      w = c - self.p
      wa = (w @ a)[..., None]*a
      wp = w - wa
      cr = np.cross(a, wp)
      return self.p + wa + np.cos(uu)[..., None]*wp + np.sin(uu)[..., None]*cr
    if k in ('bspline', 'bezier'):
      uu, vv = np.broadcast_arrays(u, v)
      out = np.empty(uu.shape+(3,))
      for idx in np.ndindex(uu.shape):
        out[idx] = self._eval_spline(float(uu[idx]), float(vv[idx]))
      return out
    raise BRepError(f'cannot evaluate surface kind {k}')

  def _eval_spline(self, u, v):
    s = self.spline
    if s['uperiod'] is not None:
      u = s['uknots'][0] + (u-s['uknots'][0]) % s['uperiod']
    if s['vperiod'] is not None:
      v = s['vknots'][0] + (v-s['vknots'][0]) % s['vperiod']
    ctrl = s['ctrl']   # [nu, nv, 4]
    # evaluate along v for every u-row, then along u
    rows = np.array([_de_boor(s['vflat'], s['vdeg'], ctrl[i], v) for i in range(ctrl.shape[0])])
    p = _de_boor(s['uflat'], s['udeg'], rows, u)
    return p[:3]/p[3]


def _flat_knots(knots, mults, degree, periodic, npoles):
  knots = np.asarray(knots, dtype=float)
  mults = np.asarray(mults, dtype=int)
  flat = np.repeat(knots, mults)
  if not periodic:
    return flat, None, np.arange(npoles)
  period = knots[-1]-knots[0]
  m0 = mults[0]
  need = degree+1-m0
  left = (flat[-m0-need:-m0]-period) if need > 0 else np.array([])
  right = (flat[m0:m0+need]+period) if need > 0 else np.array([])
  flat2 = np.concatenate([left, flat, right])
  nflat = len(flat2)-degree-1
  return flat2, period, np.arange(nflat) % npoles


def _read_surface(tk):
  typ = tk.int()
  if 1 <= typ <= 5:
    p, n, dx, dy = tk.reals(3), tk.reals(3), tk.reals(3), tk.reals(3)
    if typ == 1:
      return Surface('plane', p=p, n=n, dx=dx, dy=dy)
    if typ == 2:
      return Surface('cylinder', p=p, n=n, dx=dx, dy=dy, r=tk.real())
    if typ == 3:
      return Surface('cone', p=p, n=n, dx=dx, dy=dy, r=tk.real(), angle=tk.real())
    if typ == 4:
      return Surface('sphere', p=p, n=n, dx=dx, dy=dy, r=tk.real())
    return Surface('torus', p=p, n=n, dx=dx, dy=dy, r=tk.real(), r2=tk.real())
  if typ == 6:
    d = tk.reals(3)
    return Surface('extrusion', d=d, curve=_read_curve3d(tk))
  if typ == 7:
    p, d = tk.reals(3), tk.reals(3)
    return Surface('revolution', p=p, d=d, curve=_read_curve3d(tk))
  if typ == 8:
    urat, vrat, udeg, vdeg = tk.int(), tk.int(), tk.int(), tk.int()
    rational = bool(urat or vrat)
    ctrl = np.empty((udeg+1, vdeg+1, 4))
    for i in range(udeg+1):
      for j in range(vdeg+1):
        xyz = tk.reals(3)
        w = tk.real() if rational else 1.0
        ctrl[i, j, :3], ctrl[i, j, 3] = xyz*w, w
    uflat = np.repeat([0.0, 1.0], [udeg+1, udeg+1])
    vflat = np.repeat([0.0, 1.0], [vdeg+1, vdeg+1])
    return Surface('bezier', spline=dict(ctrl=ctrl, udeg=udeg, vdeg=vdeg, uflat=uflat, vflat=vflat,
                                         uknots=np.array([0., 1.]), vknots=np.array([0., 1.]),
                                         uperiod=None, vperiod=None))
  if typ == 9:
    urat, vrat, uper, vper = tk.int(), tk.int(), tk.int(), tk.int()
    udeg, vdeg, nup, nvp, nuk, nvk = (tk.int() for _ in range(6))
    rational = bool(urat or vrat)
    ctrl = np.empty((nup, nvp, 4))
    for i in range(nup):
      for j in range(nvp):
        xyz = tk.reals(3)
        w = tk.real() if rational else 1.0
        ctrl[i, j, :3], ctrl[i, j, 3] = xyz*w, w
    uk, um, vk, vm = [], [], [], []
    for _ in range(nuk):
      uk.append(tk.real()); um.append(tk.int())
    for _ in range(nvk):
      vk.append(tk.real()); vm.append(tk.int())
    uflat, uperiod, uidx = _flat_knots(uk, um, udeg, bool(uper), nup)
    vflat, vperiod, vidx = _flat_knots(vk, vm, vdeg, bool(vper), nvp)
    ctrl = ctrl[uidx][:, vidx]
    return Surface('bspline', spline=dict(ctrl=ctrl, udeg=udeg, vdeg=vdeg, uflat=uflat, vflat=vflat,
                                          uknots=np.array(uk), vknots=np.array(vk),
                                          uperiod=uperiod, vperiod=vperiod))
  if typ == 10:
    u1, u2, v1, v2 = tk.real(), tk.real(), tk.real(), tk.real()
    basis = _read_surface(tk)
    if basis.is_elementary():
      return basis                       # same parametrisation, trimming comes from the wires
    return Surface('rtrimmed', basis=basis, bounds=(u1, u2, v1, v2))
  if typ == 11:
    off = tk.real()
    basis = _read_surface(tk)
    if off == 0:
      return Surface('offsetsurf_zero', basis=basis)
    return Surface('offsetsurf', basis=basis, offset=off)
  raise BRepError(f'unknown surface type {typ}')


# ------------------------------------------------------------------------------------------
# topology

class TShape:
  __slots__ = ('kind', 'subs', 'flags', 'tol', 'point', 'reps', 'degenerated', 'surface', 'loc',
               'natural_restriction', 'curve3d')

  def __init__(self, kind):
    self.kind = kind
    self.subs = []      # (orientation char, TShape index 0-based, location index)
    self.reps = []
    self.curve3d = None


class PCurveRep:
  __slots__ = ('curve', 'curve2', 'surface', 'loc', 'first', 'last')


class FaceInstance:
  '''One face reached by walking the shape tree.'''
  __slots__ = ('surface', 'surface_index', 'transform', 'reversed', 'loops', 'shell_key', 'tshape_index',
               'tolerance')


class BRepShape:
  def __init__(self, text):
    self.version = 1
    m = re.search(r'CASCADE Topology V(\d)', text)
    if m:
      self.version = int(m.group(1))
    toks = text.split()
    # locate sections
    pos = {}
    for i, t in enumerate(toks):
      if t in _SECTION_KEYS and t not in pos and i+1 < len(toks) and toks[i+1].isdigit():
        pos[t] = i
    for key in ('Locations', 'Curve2ds', 'Surfaces', 'TShapes'):
      if key not in pos:
        raise BRepError(f'section {key} missing in BRep text')
    self._read_locations(_Tokens(toks, pos['Locations']+1))
    self._read_curve2ds(_Tokens(toks, pos['Curve2ds']+1))
    self._read_curves3d(_Tokens(toks, pos['Curves']+1) if 'Curves' in pos else None)
    self._read_surfaces(_Tokens(toks, pos['Surfaces']+1))
    self._read_tshapes(_Tokens(toks, pos['TShapes']+1))

  # -- sections
  def _read_locations(self, tk):
    n = tk.int()
    self.locations = [np.eye(4)]
    for _ in range(n):
      typ = tk.int()
      if typ == 1:
        m = np.eye(4)
        m[:3, :] = tk.reals(12).reshape(3, 4)
        self.locations.append(m)
      elif typ == 2:
        m = np.eye(4)
        while True:
          idx = tk.int()
          if idx == 0:
            break
          power = tk.int()
          base = self.locations[idx]
          if power < 0:
            base = np.linalg.inv(base)
          for _ in range(abs(power)):
            m = m @ base
        self.locations.append(m)
      else:
        raise BRepError(f'unknown location type {typ}')

  def _read_curve2ds(self, tk):
    n = tk.int()
    self.curve2ds = [None]
    for _ in range(n):
      self.curve2ds.append(_read_curve2d(tk))

  def _read_curves3d(self, tk):
    self.curves3d = [None]
    if tk is None:
      return
    n = tk.int()
    for _ in range(n):
      self.curves3d.append(_read_curve3d(tk))

  def _read_surfaces(self, tk):
    n = tk.int()
    self.surfaces = [None]
    for _ in range(n):
      self.surfaces.append(_read_surface(tk))

  def _read_tshapes(self, tk):
    n = tk.int()
    self.tshapes = []
    for _ in range(n):
      kind = tk.next()
      ts = TShape(kind)
      if kind == 'Ve':
        ts.tol = tk.real()
        ts.point = tk.reals(3)
        while True:
          tk.real()              # p1
          val = tk.int()
          if val == 1:
            tk.int()
          elif val == 2:
            tk.int(); tk.int()
          elif val == 3:
            tk.real(); tk.int()
          if val > 0:
            tk.int()             # location
          else:
            break
      elif kind == 'Ed':
        ts.tol = tk.real()
        tk.int(); tk.int()
        ts.degenerated = bool(tk.int())
        while True:
          val = tk.int()
          if val == 0:
            break
          if val == 1:
            ts.curve3d = (tk.int(), tk.int(), tk.real(), tk.real())
          elif val in (2, 3):
            rep = PCurveRep()
            rep.curve = tk.int()
            rep.curve2 = None
            if val == 3:
              tok = tk.next()
              m = re.match(r'^(\d+)(' + '|'.join(_CONT) + r')?$', tok)
              if not m:
                raise BRepError(f'bad closed-surface pcurve token {tok!r}')
              rep.curve2 = int(m.group(1))
              if m.group(2) is None:
                tk.next()        # continuity as its own token
            rep.surface = tk.int()
            rep.loc = tk.int()
            rep.first, rep.last = tk.real(), tk.real()
            if self.version == 2:
              tk.reals(8 if val == 3 else 4)   # stored UV end points
            ts.reps.append(rep)
          elif val == 4:
            tk.next(); tk.int(); tk.int(); tk.int(); tk.int()
          elif val == 5:
            tk.int(); tk.int()
          elif val in (6, 7):
            tk.int()
            if val == 7:
              tok = tk.next()
              if re.match(r'^\d+$', tok):
                tk.next()
            tk.int(); tk.int()
          else:
            raise BRepError(f'unknown edge representation {val}')
      elif kind == 'Fa':
        ts.natural_restriction = bool(tk.int())
        ts.tol = tk.real()
        ts.surface = tk.int()
        ts.loc = tk.int()
        if tk.peek() == '2' :
          tk.int(); tk.int()     # triangulation reference
      elif kind in ('Wi', 'Sh', 'So', 'CS', 'Co'):
        pass
      else:
        raise BRepError(f'unknown TShape kind {kind!r}')
      ts.flags = tk.next()
      if not re.match(r'^[01]{7}$', ts.flags):
        raise BRepError(f'bad TShape flags {ts.flags!r} after {kind}')
      while True:
        tok = tk.next()
        if tok == '*':
          break
        orient, idx = tok[0], int(tok[1:])
        loc = tk.int()
        ts.subs.append((orient, n-idx, loc))
      self.tshapes.append(ts)
    # root shape reference
    tok = tk.next()
    self.root = (tok[0], n-int(tok[1:]), tk.int())

  # -- traversal
  def faces(self):
    '''
    Walk the shape tree from the root and return a FaceInstance for every face, with accumulated
    rigid transform (incl. the root location = the object's Placement) and composed orientation.
    '''
    out = []
    counter = [0]

    def walk(ref, parent_m, parent_rev, shell_key):
      orient, idx, loc = ref
      ts = self.tshapes[idx]
      m = parent_m @ self.locations[loc]
      rev = parent_rev ^ (orient == '-')
      if ts.kind == 'Sh':
        counter[0] += 1
        shell_key = counter[0]
      if ts.kind == 'Fa':
        fi = FaceInstance()
        fi.surface_index = ts.surface
        fi.surface = self.surfaces[ts.surface]
        fi.transform = m @ self.locations[ts.loc]
        fi.reversed = rev
        fi.shell_key = shell_key
        fi.tshape_index = idx
        fi.tolerance = ts.tol
        fi.loops = self._face_loops(ts)
        out.append(fi)
        return
      if ts.kind in ('Ed', 'Ve'):
        return
      for sub in ts.subs:
        walk(sub, m, rev, shell_key)

    walk(self.root, np.eye(4), False, None)
    return out

  def _face_loops(self, face):
    '''
    Boundary of a face as a list of wires, each a list of (curve2d, first, last) in the (u,v)
    space of the face's surface.  A pcurve representation belongs to this face when it names the
    same surface and its location equals (wire_loc*edge_loc)^-1 * face_surface_loc
    (BRep_Tool::CurveOnSurface semantics).
    '''
    tf = self.locations[face.loc]
    loops = []
    for (_, widx, wloc) in face.subs:
      wire = self.tshapes[widx]
      if wire.kind != 'Wi':
        continue
      segs = []
      for (_, eidx, eloc) in wire.subs:
        edge = self.tshapes[eidx]
        if edge.kind != 'Ed':
          continue
        need = np.linalg.inv(self.locations[wloc] @ self.locations[eloc]) @ tf
        cands = [r for r in edge.reps if r.surface == face.surface]
        match = [r for r in cands if np.allclose(self.locations[r.loc], need, atol=1e-9)]
        if not match and len(cands) == 1:
          match = cands
        if not match and edge.curve3d is not None and self.surfaces[face.surface].kind == 'plane':
          # planar face whose edges carry only 3-D curves (OCC derives such pcurves on the fly):
          # express the 3-D curve in the plane's own (u, v) frame
          cidx, cloc, first, last = edge.curve3d
          m = np.linalg.inv(tf) @ self.locations[wloc] @ self.locations[eloc] @ self.locations[cloc]
          pc = project_curve_on_plane(self.curves3d[cidx], m, self.surfaces[face.surface])
          segs.append((pc, first, last, ('3d', cidx, first, last, eidx)))
        for r in match:
          for cidx in (r.curve, r.curve2):
            if cidx is None:
              continue
            key = (cidx, r.first, r.last)
            if key not in [s[3] for s in segs]:
              segs.append((self.curve2ds[cidx], r.first, r.last, key))
      loops.append([(c, a, b) for c, a, b, _ in segs])
    return loops


class ProjectedCurve2d:
  'generic 3-D curve seen in the (u, v) frame of a plane'
  kind = 'projected'
  def __init__(self, curve, m, plane):
    self.curve, self.m, self.plane = curve, m, plane
  def eval(self, t):
    p = self.curve.eval(np.asarray(t, dtype=float))
    p = p @ self.m[:3, :3].T + self.m[:3, 3]
    w = p - self.plane.p
    return np.stack([w @ self.plane.dx, w @ self.plane.dy], axis=-1)


def project_curve_on_plane(curve, m, plane):
  '''
  3-D curve (moved by the 4x4 matrix m into the plane's coordinate system) -> curve in the plane's
  (u, v) space.  Lines and circles parallel to the plane stay exact.
  '''
  R, T = m[:3, :3], m[:3, 3]
  def uv(p):
    w = (R @ p + T) - plane.p
    return np.array([w @ plane.dx, w @ plane.dy])
  def uvdir(d):
    w = R @ d
    return np.array([w @ plane.dx, w @ plane.dy])
  if curve.kind == 'line':
    return Line2d(uv(curve.p), uvdir(curve.d))
  if curve.kind == 'circle':
    dx, dy = uvdir(curve.dx), uvdir(curve.dy)
    if abs(np.linalg.norm(dx)-1) < 1e-9 and abs(np.linalg.norm(dy)-1) < 1e-9:
      return Circle2d(uv(curve.p), dx, dy, curve.r)
  if curve.kind == 'trimmed':
    return project_curve_on_plane(curve.basis, m, plane)
  return ProjectedCurve2d(curve, m, plane)


def read_brep(text):
  if isinstance(text, bytes):
    text = text.decode('ascii', errors='replace')
  return BRepShape(text)
