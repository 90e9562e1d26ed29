'''
Stand-ins for the FreeCAD `Part` / OpenCASCADE answers that the reference's Ray.findNearestIntersection and
Ray.getNormal ask for (reference freecad_elements/ray.py:345-426,463-470), so that the reference's OWN selection
logic — find.relevantOpticalObjects, the shell / face bounding-box culls with the LINE test, the three acceptance
rules, the maxRayLength shrink, the minDist + 2 tol filter, the "prefer a group that is not the current medium"
choice — can be executed in a container without FreeCAD.  TEST INFRASTRUCTURE for tests/golden/make_traceray_golden.py;
nothing here is shipped and nothing here touches oracle/ (numpy only): the golden it produces judges the oracle.

What is restated, and from which documented behaviour:
  Part.makeLine(p1, p2)            an edge p1 -> p2 whose .Curve is the INFINITE Geom_Line through p1 along p2 - p1
  line.Curve.intersect(surface)    GeomAPI_IntCS of the infinite line with the UNTRIMMED surface -> ([Part.Point, ...], [])
  Part.Vertex(p).distToShape(s)    BRepExtrema_DistShapeShape: (minimum distance, ...) to the finite edge / to the trimmed face
  BoundBox.isInside / closestPoint / enlarge / intersect(base, dir)
                                   Base::BoundBox3: inclusive containment, component-wise clamp, in-place growth, and
                                   IsCutLine — a test against the infinite LINE (quirk Q4 of SURVEY.md), done like FreeCAD
                                   does it: cut the line with each of the six side planes and test the cut point
                                   against the other two coordinate ranges
  Surface.parameter(p)             ElSLib::Parameters of the elementary surfaces (u in [0, 2 pi))
  Face.normalAt(u, v)              unit normal of the oriented face
Surfaces are the elementary ones (plane, cylinder, cone, sphere, torus) with OCC's parametrisations; faces are a
surface + a trimming region in (u, v) (whole surface, (u, v) box, or loops of straight / circular pcurves).

Line/surface crossings are found as the real roots of the polynomial obtained by substituting the line into the
implicit equation in the surface's own frame (numpy.roots + Newton polish): a different algorithm from the
closed-form solvers of the oracle and the kernels.  The distance of a point to a trimmed face is |P - foot| when the
foot point's (u, v) lies inside the trimming region and the 3-D distance to the nearest boundary curve otherwise
(boundary curves sampled and refined by a local search, so the value is never BELOW the true distance: a point that is
really off the face by more than the tolerance is never accepted).
'''
import math

import numpy as np

from freecad_stub import Vector

PLANE, CYLINDER, CONE, SPHERE, TORUS = 1, 2, 3, 4, 5
TRIM_NONE, TRIM_UVBOX, TRIM_LOOPS = 0, 1, 2
SEG_LINE, SEG_ARC = 1, 2
TWO_PI = 2*math.pi


def _v(p):
  return np.array([p[0], p[1], p[2]], dtype=np.float64)


class Point:
  'Part.Point: what Curve.intersect returns'
  def __init__(self, p):
    self.X, self.Y, self.Z = float(p[0]), float(p[1]), float(p[2])


class BoundBox:
  def __init__(self, lo, hi):
    self.lo, self.hi = np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)

  def enlarge(self, d):
    self.lo = self.lo - float(d)
    self.hi = self.hi + float(d)

  def isInside(self, p):
    p = _v(p)
    return bool(np.all(p >= self.lo) and np.all(p <= self.hi))

  def closestPoint(self, p):
    return Vector(np.minimum(np.maximum(_v(p), self.lo), self.hi))

  def intersect(self, base, direction):
    'Base::BoundBox3::IsCutLine with tolerance 0'
    b, d = _v(base), _v(direction)
    centre = (self.lo+self.hi)/2
    if np.linalg.norm(np.cross(d, centre-b))/np.linalg.norm(d) > np.linalg.norm(self.hi-self.lo):
      return False
    for axis in range(3):
      if d[axis] == 0.0:
        continue                                     # the line is parallel to this pair of side planes
      others = [a for a in range(3) if a != axis]
      for side in (self.lo[axis], self.hi[axis]):
        cut = b + d*((side-b[axis])/d[axis])
        if all(self.lo[a] <= cut[a] <= self.hi[a] for a in others):
          return True
    return False


class Surface:
  'an elementary surface in its own frame (origin O, axes X, Y, Z), OCC parametrisation'
  def __init__(self, kind, O, X, Y, Z, p0, p1):
    self.kind, self.O, self.X, self.Y, self.Z, self.p0, self.p1 = int(kind), _v(O), _v(X), _v(Y), _v(Z), float(p0), float(p1)

  def local(self, P):
    w = _v(P)-self.O
    return np.array([w@self.X, w@self.Y, w@self.Z])

  def world(self, q):
    q = np.asarray(q, dtype=np.float64)
    return self.O + q[..., 0:1]*self.X + q[..., 1:2]*self.Y + q[..., 2:3]*self.Z

  # --- ElSLib::Parameters -------------------------------------------------------------------
  def parameter(self, P):
    x, y, z = self.local(P)
    ang = math.atan2(y, x) % TWO_PI
    if self.kind == PLANE:
      return (x, y)
    if self.kind == CYLINDER:
      return (ang, z)
    if self.kind == CONE:
      v = z/math.cos(self.p1)
      if self.p0 + v*math.sin(self.p1) < 0:          # second nappe: the radial direction of the parametrisation is reversed
        ang = (ang+math.pi) % TWO_PI
      return (ang, v)
    if self.kind == SPHERE:
      return (ang, math.atan2(z, math.hypot(x, y)))
    if self.kind == TORUS:
      return (ang, math.atan2(z, math.hypot(x, y)-self.p0) % TWO_PI)
    raise ValueError('surface kind')

  def value_local(self, u, v):
    u, v = np.asarray(u, dtype=np.float64), np.asarray(v, dtype=np.float64)
    cu, su = np.cos(u), np.sin(u)
    if self.kind == PLANE:
      return np.stack([u, v, np.zeros_like(u)], axis=-1)
    if self.kind == CYLINDER:
      return np.stack([self.p0*cu, self.p0*su, v+0*u], axis=-1)
    if self.kind == CONE:
      r = self.p0 + v*math.sin(self.p1)
      return np.stack([r*cu, r*su, v*math.cos(self.p1)+0*u], axis=-1)
    if self.kind == SPHERE:
      return np.stack([self.p0*np.cos(v)*cu, self.p0*np.cos(v)*su, self.p0*np.sin(v)+0*u], axis=-1)
    r = self.p0 + self.p1*np.cos(v)
    return np.stack([r*cu, r*su, self.p1*np.sin(v)+0*u], axis=-1)

  def value(self, u, v):
    return self.world(self.value_local(u, v))

  def normal_geom(self, u, v):
    'unit normal pointing radially outward (plane: +Z), written with the frame axes'
    cu, su = math.cos(u), math.sin(u)
    rad = cu*self.X + su*self.Y
    if self.kind == PLANE:
      n = self.Z
    elif self.kind == CYLINDER:
      n = rad
    elif self.kind == CONE:
      sg = 1.0 if self.p0 + v*math.sin(self.p1) >= 0 else -1.0
      n = sg*(math.cos(self.p1)*rad - sg*math.sin(self.p1)*self.Z)       # perpendicular to the generator, away from the axis
    else:                                                                # sphere, torus: along the (tube) radius
      n = math.cos(v)*rad + math.sin(v)*self.Z
    return n/np.linalg.norm(n)

  # --- GeomAPI_IntCS for an infinite line -----------------------------------------------------
  def line_parameters(self, start, direction):
    'all t with start + t*direction on the untrimmed surface (direction need not be unit)'
    s, d = self.local(start), np.array([_v(direction)@self.X, _v(direction)@self.Y, _v(direction)@self.Z])
    x, y, z = (np.array([d[i], s[i]]) for i in range(3))          # polynomials in t, highest power first
    sq = lambda p: np.polymul(p, p)
    if self.kind == PLANE:
      poly = z
    elif self.kind == SPHERE:
      poly = np.polysub(np.polyadd(np.polyadd(sq(x), sq(y)), sq(z)), [self.p0**2])
    elif self.kind == CYLINDER:
      poly = np.polysub(np.polyadd(sq(x), sq(y)), [self.p0**2])
    elif self.kind == CONE:
      rho = np.polyadd([self.p0], math.tan(self.p1)*z)
      poly = np.polysub(np.polyadd(sq(x), sq(y)), sq(rho))
    else:
      R, r = self.p0, self.p1
      m = np.polyadd(np.polyadd(np.polyadd(sq(x), sq(y)), sq(z)), [R*R-r*r])
      poly = np.polysub(sq(m), 4*R*R*np.polyadd(sq(x), sq(y)))
    scale = np.abs(poly).max()
    if scale == 0:
      return []
    lead = 0
    while lead < len(poly)-1 and abs(poly[lead]) <= 1e-14*scale:  # e.g. a line parallel to a cylinder axis
      lead += 1
    poly = poly[lead:]
    if len(poly) < 2:
      return []
    dp = np.polyder(poly)
    out = []
    for root in np.roots(poly):
      if abs(root.imag) > 1e-6*max(1.0, abs(root.real)):
        continue
      t = float(root.real)
      for _ in range(4):                                          # Newton polish on the polynomial
        f, g = np.polyval(poly, t), np.polyval(dp, t)
        if g == 0:
          break
        t -= f/g
      # a complex pair of a near-tangent line can pass the imaginary-part filter: keep only true zeros
      residual = abs(np.polyval(poly, t))
      size = np.polyval(np.abs(poly), abs(t))
      if residual <= 1e-9*size and not any(abs(t-o) <= 1e-12*max(1.0, abs(t)) for o in out):
        out.append(t)
    return sorted(out)


class Curve:
  'Geom_Line'
  def __init__(self, start, direction):
    self.start, self.direction = _v(start), _v(direction)

  def intersect(self, surface):
    ts = surface.line_parameters(self.start, self.direction)
    return [Point(self.start + t*self.direction) for t in ts], []


class LineShape:
  'Part.makeLine(p1, p2): a finite edge; .Curve is the infinite line'
  def __init__(self, p1, p2):
    self.p1, self.p2 = _v(p1), _v(p2)
    self.Curve = Curve(self.p1, self.p2-self.p1)

  def distance(self, p):
    d = self.p2-self.p1
    t = min(1.0, max(0.0, float((p-self.p1)@d/(d@d))))
    return float(np.linalg.norm(self.p1 + t*d - p))


class Face:
  '''
  A trimmed face.  rec: one row of the scene's face table (its frame already in the coordinates the shape lives in),
  segs: the trim segment table.  index = row in the scene (what the golden reports as face).
  '''
  def __init__(self, rec, segs, index):
    self.index = index
    self.Surface = Surface(rec['kind'], rec['origin'], rec['xdir'], rec['ydir'], rec['zdir'], rec['p0'], rec['p1'])
    self.trim = int(rec['trim_kind'])
    self.nsign = int(rec['nsign'])
    self.uv_min, self.uv_max = np.array(rec['uv_min'], dtype=np.float64), np.array(rec['uv_max'], dtype=np.float64)
    self.segs = [(int(s['kind']), np.array(s['a'], dtype=np.float64)) for s in segs[int(rec['seg_first']):int(rec['seg_first'])+int(rec['seg_count'])]
                 if int(s['kind']) in (SEG_LINE, SEG_ARC)] if self.trim == TRIM_LOOPS else []
    self._lo, self._hi = np.array(rec['aabb_min'], dtype=np.float64), np.array(rec['aabb_max'], dtype=np.float64)
    k = self.Surface.kind
    self.u_periodic = k != PLANE
    self.v_periodic = k == TORUS
    self._boundary = None

  @property
  def BoundBox(self):
    return BoundBox(self._lo, self._hi)             # a new object per access, like FreeCAD

  def normalAt(self, u, v):
    return Vector(self.nsign*self.Surface.normal_geom(u, v))

  # --- trimming region ------------------------------------------------------------------------
  def _window(self, u, v):
    if self.u_periodic:
      u = self.uv_min[0] + (u-self.uv_min[0]) % TWO_PI
    if self.v_periodic:
      v = self.uv_min[1] + (v-self.uv_min[1]) % TWO_PI
    return u, v

  def contains(self, u, v):
    'is (u, v) inside the trimming region (exact, no tolerance)'
    if self.trim == TRIM_NONE:
      return True
    u, v = self._window(u, v)
    if self.trim == TRIM_UVBOX:
      return self.uv_min[0] <= u <= self.uv_max[0] and self.uv_min[1] <= v <= self.uv_max[1]
    crossings = 0                                    # even-odd rule along the half line (u' > u, v)
    for kind, a in self.segs:
      if kind == SEG_LINE:
        u0, v0, u1, v1 = a[:4]
        if (v0 > v) != (v1 > v):
          if u0 + (v-v0)/(v1-v0)*(u1-u0) > u:
            crossings += 1
      else:
        cu, cv, r, a0, span = a
        h2 = r*r - (v-cv)**2
        if h2 > 0:
          for uc in (cu-math.sqrt(h2), cu+math.sqrt(h2)):
            if uc > u and (math.atan2(v-cv, uc-cu)-a0) % TWO_PI <= span:
              crossings += 1
    return crossings % 2 == 1

  def _boundary_curves(self):
    'list of (u(s), v(s)) callables over s in [0, 1] describing the boundary pcurves'
    curves = []
    if self.trim == TRIM_UVBOX:
      (u0, v0), (u1, v1) = self.uv_min, self.uv_max
      full_u = self.u_periodic and abs((u1-u0)-TWO_PI) < 1e-9
      full_v = self.v_periodic and abs((v1-v0)-TWO_PI) < 1e-9
      if not full_v:
        curves += [lambda s, v=v0: (u0+s*(u1-u0), v+0*s), lambda s, v=v1: (u0+s*(u1-u0), v+0*s)]
      if not full_u:
        curves += [lambda s, u=u0: (u+0*s, v0+s*(v1-v0)), lambda s, u=u1: (u+0*s, v0+s*(v1-v0))]
    for kind, a in self.segs:
      if kind == SEG_LINE:
        curves.append(lambda s, a=a: (a[0]+s*(a[2]-a[0]), a[1]+s*(a[3]-a[1])))
      else:
        curves.append(lambda s, a=a: (a[0]+a[2]*np.cos(a[3]+s*a[4]), a[1]+a[2]*np.sin(a[3]+s*a[4])))
    return curves

  def boundary_distance(self, P, samples=1025):
    if self._boundary is None:
      s = np.linspace(0.0, 1.0, samples)
      self._boundary = [(c, s, self.Surface.value(*c(s))) for c in self._boundary_curves()]
    best = math.inf
    for c, s, pts in self._boundary:
      d = np.linalg.norm(pts-P, axis=1)
      i = int(np.argmin(d))
      lo, hi = s[max(i-1, 0)], s[min(i+1, len(s)-1)]
      f = lambda t: float(np.linalg.norm(self.Surface.value(*c(np.array(t)))-P))
      for _ in range(60):                            # golden-section search between the neighbours of the best sample
        m1, m2 = lo+(hi-lo)*0.381966011250105, lo+(hi-lo)*0.618033988749895
        if f(m1) < f(m2):
          hi = m2
        else:
          lo = m1
      best = min(best, float(d[i]), f((lo+hi)/2))
    return best

  def distance(self, P):
    'BRepExtrema distance of a point to the trimmed face'
    u, v = self.Surface.parameter(P)
    foot = self.Surface.value(u, v)
    if self.contains(u, v):
      return float(np.linalg.norm(foot-P))
    return self.boundary_distance(P)


class Shell:
  def __init__(self, faces, lo, hi):
    self.Faces, self._lo, self._hi = list(faces), np.array(lo, dtype=np.float64), np.array(hi, dtype=np.float64)

  @property
  def BoundBox(self):
    return BoundBox(self._lo, self._hi)


class Shape:
  'group.Shape: a compound of shells'
  def __init__(self, shells):
    self.Shells = list(shells)
    self.Faces = [f for s in self.Shells for f in s.Faces]

  @property
  def BoundBox(self):
    return BoundBox(np.min([s._lo for s in self.Shells], axis=0), np.max([s._hi for s in self.Shells], axis=0))


class Vertex:
  'Part.Vertex(point)'
  def __init__(self, p):
    self.p = np.array([p.X, p.Y, p.Z]) if hasattr(p, 'X') else _v(p)

  def distToShape(self, shape):
    return (shape.distance(self.p), [], [])


def group_shape(scene, group, faces=None):
  'Shape of one optical group of a flat scene (tests: the scene arrays of scene_export.scene.Scene)'
  faces = scene.faces if faces is None else faces
  shells = []
  for sh in scene.shells:
    if int(sh['group']) != group:
      continue
    first, count = int(sh['face_first']), int(sh['face_count'])
    fs = [Face(faces[i], scene.segs, i) for i in range(first, first+count)]
    lo = np.min([f._lo for f in fs], axis=0)
    hi = np.max([f._hi for f in fs], axis=0)
    shells.append(Shell(fs, lo, hi))
  return Shape(shells)


def install(part_module):
  'fills the stand-in `Part` module registered by freecad_stub.install()'
  part_module.makeLine = lambda p1, p2: LineShape(p1, p2)
  part_module.Vertex = Vertex
  part_module.Point = Point
