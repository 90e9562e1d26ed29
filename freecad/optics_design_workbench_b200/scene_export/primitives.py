'''
Procedural shapes expressed as brep.FaceInstance lists, i.e. in exactly the form the BRep reader
hands to the scene builder.  Used to build scenes without a FreeCAD document (tests, scripted
set-ups): every shape goes through the same face_record() classification and trimming code as an
imported Part.Shape would.

Conventions follow OCC's primitives (BRepPrimAPI): a solid's faces carry outward normals; face
orientation `reversed` flips the geometric normal du x dv.
'''

import numpy as np

from .brep import FaceInstance, Surface, Line2d, Circle2d

TWO_PI = 2*np.pi
_X, _Y, _Z = np.eye(3)


def _face(surface, loops, reversed_=False, shell_key=1, transform=None):
  fi = FaceInstance()
  fi.surface = surface
  fi.surface_index = 0
  fi.transform = np.eye(4) if transform is None else np.asarray(transform, dtype=float)
  fi.reversed = bool(reversed_)
  fi.loops = loops
  fi.shell_key = shell_key
  fi.tshape_index = -1
  fi.tolerance = 1e-7
  return fi


def _rect_loop(u0, u1, v0, v1):
  return [(Line2d(np.array([u0, v0]), np.array([1.0, 0.0])), 0.0, u1-u0),
          (Line2d(np.array([u1, v0]), np.array([0.0, 1.0])), 0.0, v1-v0),
          (Line2d(np.array([u0, v1]), np.array([1.0, 0.0])), 0.0, u1-u0),
          (Line2d(np.array([u0, v0]), np.array([0.0, 1.0])), 0.0, v1-v0)]


def _circle_loop(r, cu=0.0, cv=0.0):
  return [(Circle2d(np.array([cu, cv]), np.array([1.0, 0.0]), np.array([0.0, 1.0]), r), 0.0, TWO_PI)]


def plane_surface(origin, xdir, ydir):
  xdir, ydir = np.asarray(xdir, float), np.asarray(ydir, float)
  return Surface('plane', p=np.asarray(origin, float), n=np.cross(xdir, ydir), dx=xdir, dy=ydir)


def rectangle(lx, ly, shell_key=None):
  'open rectangular face in the z=0 plane, normal +z (a detector sheet)'
  return [_face(plane_surface((0, 0, 0), _X, _Y), [_rect_loop(0, lx, 0, ly)], shell_key=shell_key)]


def disc(r, shell_key=None):
  'open circular face in the z=0 plane, normal +z'
  return [_face(plane_surface((0, 0, 0), _X, _Y), [_circle_loop(r)], shell_key=shell_key)]


def box(lx, ly, lz):
  'solid box [0,lx]x[0,ly]x[0,lz], 6 planar faces with outward normals'
  f = []
  # (origin, xdir, ydir, u-extent, v-extent); normal = xdir x ydir
  specs = [((0, 0, 0), _Y, _X, ly, lx),      # bottom  -z
           ((0, 0, lz), _X, _Y, lx, ly),     # top     +z
           ((0, 0, 0), _X, _Z, lx, lz),      # front   -y
           ((0, ly, 0), _Z, _X, lz, lx),     # back    +y
           ((0, 0, 0), _Z, _Y, lz, ly),      # left    -x
           ((lx, 0, 0), _Y, _Z, ly, lz)]     # right   +x
  for o, xd, yd, ue, ve in specs:
    f.append(_face(plane_surface(o, xd, yd), [_rect_loop(0, ue, 0, ve)]))
  return f


def _std(kind, **kw):
  return Surface(kind, p=np.zeros(3), n=_Z.copy(), dx=_X.copy(), dy=_Y.copy(), **kw)


def sphere(radius):
  'full sphere centred at the origin: one face, boundary = seam + two degenerate pole edges'
  return [_face(_std('sphere', r=radius), [_rect_loop(0, TWO_PI, -np.pi/2, np.pi/2)])]


def torus(R, r):
  return [_face(_std('torus', r=R, r2=r), [_rect_loop(0, TWO_PI, 0, TWO_PI)])]


def cylinder(radius, height):
  'solid cylinder along +z from z=0 to z=height'
  lateral = _face(_std('cylinder', r=radius), [_rect_loop(0, TWO_PI, 0, height)])
  bottom = _face(plane_surface((0, 0, 0), _X, _Y), [_circle_loop(radius)], reversed_=True)
  top = _face(plane_surface((0, 0, height), _X, _Y), [_circle_loop(radius)])
  return [lateral, bottom, top]


def cone(r1, r2, height):
  'solid truncated cone along +z: radius r1 at z=0, r2 at z=height'
  alpha = np.arctan2(r2-r1, height)
  slant = height/np.cos(alpha)
  faces = [_face(_std('cone', r=r1, angle=alpha), [_rect_loop(0, TWO_PI, 0, slant)])]
  if r1 > 0:
    faces.append(_face(plane_surface((0, 0, 0), _X, _Y), [_circle_loop(r1)], reversed_=True))
  if r2 > 0:
    faces.append(_face(plane_surface((0, 0, height), _X, _Y), [_circle_loop(r2)]))
  return faces


def plano_convex_lens(R, aperture_radius, edge_thickness=0.0):
  '''
  Flat face at z=0 (normal -z), optional cylindrical rim of height edge_thickness, spherical cap of
  radius R bulging towards +z.  Same construction as the benchmark lens (sphere ∩ cylinder).
  '''
  sag_centre = edge_thickness - np.sqrt(R*R - aperture_radius**2)      # z of the sphere centre
  v0 = np.arcsin((edge_thickness - sag_centre)/R)
  cap_surface = Surface('sphere', p=np.array([0, 0, sag_centre]), n=_Z.copy(), dx=_X.copy(), dy=_Y.copy(), r=R)
  faces = [_face(cap_surface, [_rect_loop(0, TWO_PI, v0, np.pi/2)]),
           _face(plane_surface((0, 0, 0), _X, _Y), [_circle_loop(aperture_radius)], reversed_=True)]
  if edge_thickness > 0:
    faces.append(_face(_std('cylinder', r=aperture_radius), [_rect_loop(0, TWO_PI, 0, edge_thickness)]))
  return faces


def conicoid_surface(c, k, vertex=(0, 0, 0), poly=None):
  '''
  conic of revolution about +z with its vertex at `vertex`: sag c rho^2 / (1 + sqrt(1 - (1+k) c^2 rho^2)) (include/odw.h);
  poly = even-asphere coefficients of rho^4, rho^6, ... rho^12 added to it (the lens-catalogue form of an asphere)
  '''
  return Surface('conicoid', p=np.asarray(vertex, float), n=_Z.copy(), dx=_X.copy(), dy=_Y.copy(), c=float(c), k=float(k),
                 poly=None if poly is None else [float(x) for x in poly])


def conic_dish(c, k, aperture_radius, inner_radius=0.0, reversed_=False, poly=None):
  '''
  Open single-face shell: the conic of revolution z = sag(rho) for inner_radius <= rho <= aperture_radius (a parabolic
  mirror for k = -1: focal length 1/(2c), focus at z = 1/(2c)).  Geometric normal c rho - q z: away from the focus side.
  '''
  return [_face(conicoid_surface(c, k, poly=poly), [_rect_loop(0, TWO_PI, inner_radius, aperture_radius)], reversed_=reversed_, shell_key=None)]


def revolved_parabola_dish(focal, aperture_radius):
  '''
  The same paraboloid the way OCC writes it: surface of revolution (BRep surface type 7) of a Geom_Parabola about its own
  axis, trimmed to 0 <= t <= aperture_radius.  Exercises the revolution -> conicoid recognition of scene.face_record.
  '''
  from .brep import Curve3d
  par = Curve3d('parabola', p=np.zeros(3), n=_Y.copy(), dx=_Z.copy(), dy=_X.copy(), f=float(focal))
  surf = Surface('revolution', p=np.zeros(3), d=_Z.copy(), curve=par)
  return [_face(surf, [_rect_loop(0, TWO_PI, 0.0, aperture_radius)], shell_key=None)]


def revolved_spline_dish(sag, aperture_radius, n_points=41, reversed_=False):
  '''
  An aspheric surface the way FreeCAD users model it: a cubic B-spline through n_points points (rho, sag(rho)) of the lens
  formula in the x-z plane, revolved about the z axis (BRep surface type 7 with a B-spline generatrix), trimmed to the whole
  curve.  Exercises the revolution -> even-asphere fit of scene.face_record.
  '''
  import scipy.interpolate
  from .brep import Curve3d, BSplineCurve
  rho = np.linspace(0.0, aperture_radius, n_points)
  pts = np.column_stack([rho, np.zeros_like(rho), sag(rho)])
  sp = scipy.interpolate.make_interp_spline(rho, pts, k=3)           # parameter = rho at the data points (not in general in between)
  knots, mults = np.unique(sp.t, return_counts=True)
  curve = Curve3d('bspline', spline=BSplineCurve(3, sp.c, None, knots, mults, False))
  surf = Surface('revolution', p=np.zeros(3), d=_Z.copy(), curve=curve)
  return [_face(surf, [_rect_loop(0, TWO_PI, float(knots[0]), float(knots[-1]))], reversed_=reversed_, shell_key=None)]


# ------------------------------------------------------------------------------------------
# rigid transforms

def translation(x, y, z):
  m = np.eye(4)
  m[:3, 3] = (x, y, z)
  return m


def rotation(axis, angle):
  'Rodrigues rotation matrix (4x4) about `axis` by `angle` radians'
  a = np.asarray(axis, float)
  a = a/np.linalg.norm(a)
  K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
  m = np.eye(4)
  m[:3, :3] = np.eye(3) + np.sin(angle)*K + (1-np.cos(angle))*(K @ K)
  return m
