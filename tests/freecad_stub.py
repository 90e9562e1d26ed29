'''
Stand-ins for the three FreeCAD Base types the reference's ray code computes with (Vector, Rotation, Matrix), so
that the reference's OWN Python (ray.py, point_source.py, optical_group.py) can be executed in a container without
FreeCAD.  Test infrastructure for tests/golden/make_traceray_golden.py only; nothing here is shipped.

Semantics restated from FreeCAD's documented Base API (src/Base/VectorPy, RotationPy, MatrixPy):
  Vector * Vector   dot product (float);  Vector * number, number * Vector, Vector / number   scaling
  Vector.cross / .dot / .Length / [i] / .x .y .z,  unary minus, + and -
  Rotation(axis: Vector, angle_in_degrees)  right-handed rotation about the axis;  Rotation * Rotation composes
  (the right factor is applied first);  Rotation * Vector rotates the vector
  Matrix * Vector   the affine map applied to a point (rotation + translation)
'''
import math
import numbers
import sys
import types

import numpy as np


class Vector:
  __array_ufunc__ = None            # numpy scalars defer to __rmul__ instead of broadcasting over the components

  def __init__(self, x=0.0, y=0.0, z=0.0):
    if not isinstance(x, numbers.Real):
      x, y, z = (float(c) for c in x)
    self.x, self.y, self.z = float(x), float(y), float(z)

  def __iter__(self):
    return iter((self.x, self.y, self.z))

  def __len__(self):
    return 3

  def __getitem__(self, i):
    return (self.x, self.y, self.z)[i]

  def __add__(self, o):
    return Vector(self.x+o.x, self.y+o.y, self.z+o.z)

  def __sub__(self, o):
    return Vector(self.x-o.x, self.y-o.y, self.z-o.z)

  def __neg__(self):
    return Vector(-self.x, -self.y, -self.z)

  def __mul__(self, o):
    if isinstance(o, Vector):
      return self.x*o.x + self.y*o.y + self.z*o.z
    if isinstance(o, numbers.Real):
      return Vector(self.x*float(o), self.y*float(o), self.z*float(o))
    return NotImplemented

  __rmul__ = __mul__

  def __truediv__(self, o):
    return Vector(self.x/float(o), self.y/float(o), self.z/float(o))

  def dot(self, o):
    return self*o

  def cross(self, o):
    return Vector(self.y*o.z - self.z*o.y, self.z*o.x - self.x*o.z, self.x*o.y - self.y*o.x)

  @property
  def Length(self):
    return math.sqrt(self.x*self.x + self.y*self.y + self.z*self.z)

  def __eq__(self, o):
    return isinstance(o, Vector) and (self.x, self.y, self.z) == (o.x, o.y, o.z)

  def __hash__(self):
    return hash((self.x, self.y, self.z))

  def __repr__(self):
    return f'Vector ({self.x}, {self.y}, {self.z})'


class Rotation:
  def __init__(self, axis=None, angle=0.0, _m=None):
    if _m is not None:
      self.m = _m
      return
    a = np.array(list(axis), dtype=np.float64)
    a = a/np.linalg.norm(a)
    t = math.radians(float(angle))
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    self.m = np.eye(3) + math.sin(t)*K + (1-math.cos(t))*(K@K)          # Rodrigues

  def __mul__(self, o):
    if isinstance(o, Rotation):
      return Rotation(_m=self.m@o.m)
    if isinstance(o, Vector):
      return Vector(self.m@np.array(list(o)))
    return NotImplemented


class Matrix:
  def __init__(self, a=None):
    self.a = np.eye(4) if a is None else np.array(a, dtype=np.float64).reshape(4, 4)

  def __mul__(self, o):
    if isinstance(o, Vector):
      return Vector(self.a[:3, :3]@np.array(list(o)) + self.a[:3, 3])
    if isinstance(o, Matrix):
      return Matrix(self.a@o.a)
    return NotImplemented

  def inverse(self):
    return Matrix(np.linalg.inv(self.a))


def install():
  'registers FreeCAD / FreeCADGui / Part modules made of the stand-ins (before the reference modules are imported)'
  app = types.ModuleType('FreeCAD')
  app.Vector, app.Rotation, app.Matrix, app.GuiUp = Vector, Rotation, Matrix, False
  gui = types.ModuleType('FreeCADGui')
  part = types.ModuleType('Part')
  sys.modules.update({'FreeCAD': app, 'FreeCADGui': gui, 'Part': part})
  return app
