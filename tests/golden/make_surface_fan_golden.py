'''
Generates tests/golden/surface_fan_golden.npz by running the reference's OWN surface-source fan mode
  SurfaceSourceProxy._makeSurfaceGrid   (freecad_elements/surface_source.py:122-267)
  SurfaceSourceProxy._generateRays(mode='fans') + _makeRay   (:418-517, :85-111)
imported unmodified from /root/reference under the FreeCAD stand-ins of tests/freecad_stub.py.  The questions the
reference asks OpenCASCADE about a face (ParameterRange, valueAt, derivative1At, normalAt, Area, and
Part.Vertex(p).distToShape(face) for points it has just evaluated ON the surface) are answered by
freecad_elements.surface_source.FaceEvaluator — the same closed forms the engine's host side uses — so the golden pins
the restated grid algorithm (pass structure, rounding, thinning, face skipping, ray construction), not the geometry.
Run here only:  python tests/golden/make_surface_fan_golden.py
'''
import os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

import freecad_stub
app = freecad_stub.install()
from freecad_stub import Vector, Matrix
from make_fan_golden import load_reference_point_source
import surface_fan_cases as cases
from freecad.optics_design_workbench_b200.freecad_elements import surface_source as ss


class PointOnFace(Vector):
  'valueAt result that remembers its parameters (for the distToShape stand-in)'
  def __init__(self, p, uv):
    super().__init__(p)
    self.uv = uv


class FaceStub:
  def __init__(self, evaluator, tol):
    self.ev, self.tol = evaluator, tol
    self.ParameterRange = tuple(evaluator.parameter_range())
    self.Area = evaluator.area
  def valueAt(self, u, v):
    return PointOnFace(self.ev.value_at(u, v), (u, v))
  def derivative1At(self, u, v):
    du, dv = self.ev.derivative1_at(u, v)
    return Vector(du), Vector(dv)
  def normalAt(self, u, v):
    return Vector(self.ev.normal_at(u, v))


class Vertex:
  def __init__(self, p):
    self.p = p
  def distToShape(self, face):
    return (0.0 if face.ev.on_face(*self.p.uv, face.tol) else 1e30, [], [])


def main():
  sys.modules['Part'].Vertex = Vertex
  ps = load_reference_point_source()
  from odw_ref.freecad_elements import surface_source as ref
  ref.keepGuiResponsiveAndRaiseIfSimulationDone = lambda *a, **k: None
  ref.Part = sys.modules['Part']
  identity = Matrix()
  ref.allCoordinateTransformMatrices = lambda part: [(identity, identity, identity, identity)]
  out = {}
  for name, (build, tol, counts) in cases.GRID_CASES.items():
    emit = build()
    proxy = ref.SurfaceSourceProxy.__new__(ref.SurfaceSourceProxy)
    proxy._getDistTol = lambda distTol=None, _t=tol: _t
    face = FaceStub(ss.FaceEvaluator(emit.faces[0], emit.segs), tol)
    for n in counts:
      grid = proxy._makeSurfaceGrid(None, face, n)
      out[f'{name}@{n}/uv'] = np.array([g[0] for g in grid]).reshape(-1, 2)
      out[f'{name}@{n}/points'] = np.array([tuple(g[1]) for g in grid]).reshape(-1, 3)
      out[f'{name}@{n}/du'] = np.array([tuple(g[2][0]) for g in grid]).reshape(-1, 3)
      out[f'{name}@{n}/dv'] = np.array([tuple(g[2][1]) for g in grid]).reshape(-1, 3)
      print(name, n, '->', len(grid), 'grid points', flush=True)
  for name, (build, tol, count) in cases.SOURCE_CASES.items():
    emit = build()
    proxy = ref.SurfaceSourceProxy.__new__(ref.SurfaceSourceProxy)
    proxy._getDistTol = lambda distTol=None, _t=tol: _t
    faces = {f'Face{i+1}': FaceStub(ss.FaceEvaluator(emit.faces[i], emit.segs), tol) for i in range(len(emit.faces))}
    proxy._cachedSelectedFace = lambda part, attr: faces[attr]
    obj = types.SimpleNamespace(Name='OpticalSurfaceSource', Label='OpticalSurfaceSource', Wavelength=500.0, FanModeRayCount=count,
                                Placement=types.SimpleNamespace(isIdentity=lambda: True),
                                ActiveSurfaces=[(types.SimpleNamespace(Label='Body', Name='Body'), list(faces))])
    import warnings
    with warnings.catch_warnings():
      warnings.simplefilter('ignore')
      try:
        rays = list(proxy._generateRays(obj, mode='fans'))
      except NameError as e:                    # the reference's own warning text references an undefined name (rayCount)
        print(name, 'reference raised', repr(e))
        continue
    out[f'{name}/origins'] = np.array([tuple(r.initPoint) for r in rays]).reshape(-1, 3)
    out[f'{name}/directions'] = np.array([tuple(r.initDirection) for r in rays]).reshape(-1, 3)
    print(name, count, '->', len(rays), 'rays', flush=True)
  path = os.path.join(HERE, 'surface_fan_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
