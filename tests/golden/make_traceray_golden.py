'''
Generates tests/golden/traceray_golden.npz by running the reference's OWN Python for the hot path
  PointSourceProxy._makeRay            (freecad_elements/point_source.py:411-460)
  Ray.traceRay / getNormal / mirror / snellsLaw / lineGrating   (freecad_elements/ray.py:36-281,455-539)
  OpticalGroupProxy.onRayHit / applyStochasticRayCorrections   (freecad_elements/optical_group.py:206-209,279-323;
                                        its (theta, phi) draws come from the engine's Philox stream: numpy's process-seeded
                                        RNG of the reference is not reproducible — the rotation formula is the reference's)
imported unmodified from /root/reference in the build container.  FreeCAD is absent, so
  * FreeCAD.Vector / Rotation / Matrix are the stand-ins of tests/freecad_stub.py, and
  * the two questions the reference asks OpenCASCADE — Ray.findNearestIntersection (ray.py:290-452) and
    Surface.parameter / Face.normalAt inside getNormal (ray.py:463-466) — are answered by the oracle's geometry
    (oracle_find_nearest / oracle_face_normal).
Everything else — the bounce loop state machine, power / medium / sequence-index bookkeeping, the maxIntersections and
powerTol exits, isEntering, the interaction formulas, what is handed to the result store — is the reference's code.
The golden therefore pins the oracle's (and the CUDA kernel's) restatement of ray.py:36-281 and point_source.py:411-460
GIVEN the geometry answers; the geometry itself stays anchored on the hand-derived known answers.
Run here only (/root/reference does not exist on the GPU box):  python tests/golden/make_traceray_golden.py
'''
import os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

import freecad_stub
freecad_stub.install()
from freecad_stub import Vector, Matrix
from make_fan_golden import load_reference_point_source
import traceray_cases as cases
from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.scene_export.scene import OPTICAL_TYPES, GRATING_TYPES
from oracle import Oracle


class Obj:
  'a document object: plain attributes, hashable by identity (raytracing_cache keys on the object)'
  def __init__(self, **kw):
    self.__dict__.update(kw)


class Store:
  'what SimulationResults.addRayHit receives (results_store.py:641-648)'
  def __init__(self):
    self.rows = []
  def addRayHit(self, source, obj, point, direction, power, isEntering, metadata):
    self.rows.append((obj.group_index, tuple(point), tuple(direction), float(power), bool(isEntering)))


def reference_modules():
  ps = load_reference_point_source()                 # imports point_source -> ray, optical_group, common with the stubs in place
  ray, og = ps.ray, sys.modules['odw_ref.freecad_elements.optical_group']
  ray.keepGuiResponsiveAndRaiseIfSimulationDone = lambda *a, **k: None
  ps.keepGuiResponsiveAndRaiseIfSimulationDone = lambda *a, **k: None
  return ps, ray, og


def make_objects(og, scene):
  objs = []
  for i, (name, g) in enumerate(zip(scene.group_names, scene.groups)):
    proxy = og.OpticalGroupProxy.__new__(og.OpticalGroupProxy)
    o = Obj(Name=name, Label=name, group_index=i, Proxy=proxy, ViewObject=None,
            OpticalType=OPTICAL_TYPES[int(g['optical_type'])], RefractiveIndex=float(g['refractive_index']),
            Reflectivity=float(g['reflectivity']), AbsorptionLength='inf', RecordHits=True,
            ReflectedProbabilityDensity='', RefractedProbabilityDensity='', RayModificationProbabilityDensity='',
            GratingType=GRATING_TYPES[int(g['grating_type'])], GratingLinesPerMillimeter=float(g['grating_lines_per_mm']),
            GratingDiffractionOrder=float(g['grating_order']), GratingLinesOrientation=Vector(g['grating_orientation']))
    objs.append(o)
  return objs


def run_case(ray_mod, oracle, scene, objs, cfg, rays, light, ignored=()):
  '''
  rays: list of reference Ray objects.  Returns per-ray segment lists and the hits handed to the store.
  '''
  sa = _abi.SceneArgs(scene)
  identity = Matrix()

  class Face:
    'the (gpM, gpMi, face) triple of ray.py:455-480 with an analytic Surface.parameter / normalAt'
    def __init__(self, index):
      self.index, self.Surface, self._n = index, self, None
    def parameter(self, point):
      uv, n = oracle.face_normal(sa, self.index, list(point))
      self._n = Vector(n)
      return tuple(uv)
    def normalAt(self, u, v):
      return self._n

  state = dict(ray=0, bounce=-1)

  class Draw:
    'stands in for the VectorRandomVariable of a surface density: compile() is a no-op (no per-hit parameters in these cases)'
    def __init__(self, group, which):
      self.group, self.which = group, which
    def compile(self, **kw):
      pass
    def draw(self):
      return oracle.scatter_draw(sa, self.group, self.which, cfg.cfg.scatter_seed, 0, state['ray'], state['bounce'])

  for o in objs:
    def get_vrv(obj, kind, _o=o):
      which = 0 if kind in ('reflect', 'refract') else 1
      return Draw(_o.group_index, which) if oracle.scatter_draw(sa, _o.group_index, which, 0, 0, 0, 0) else False
    o.Proxy._getVrv = get_vrv

  class TracedRay(ray_mod.Ray):
    def findNearestIntersection(self, start, direction, currentMedium, maxRayLength, distTol=None, sequenceIndex=None):
      state['bounce'] += 1
      fi, P = oracle.find_nearest(sa, cfg, list(start), list(direction), -1 if currentMedium is None else currentMedium.group_index,
                                  maxRayLength, sequenceIndex, ignored)
      if fi < 0:
        return None
      return objs[int(scene.faces[fi]['group'])], (identity, identity, Face(fi)), Vector(P)

  settings = Obj(MaxRayLength=cfg.cfg.max_ray_length, MaxIntersections=cfg.cfg.max_intersections)
  ray_mod.find = types.SimpleNamespace(activeSimulationSettings=lambda: settings)
  seg_p1, seg_p2, seg_power, seg_medium, seg_off = [], [], [], [], [0]
  hit_ray, hit_rows = [], []
  for i, r in enumerate(rays):
    r.__class__ = TracedRay
    state.update(ray=i, bounce=-1)
    store = Store()
    for (p1, p2), power, medium, _color in r.traceRay(store=store):
      seg_p1.append(tuple(p1)); seg_p2.append(tuple(p2)); seg_power.append(float(power))
      seg_medium.append(-1 if medium is None else medium.group_index)
    seg_off.append(len(seg_p1))
    hit_ray.extend([i]*len(store.rows))
    hit_rows.extend(store.rows)
  return dict(seg_p1=np.array(seg_p1).reshape(-1, 3), seg_p2=np.array(seg_p2).reshape(-1, 3), seg_power=np.array(seg_power),
              seg_medium=np.array(seg_medium, dtype=np.int32), seg_offsets=np.array(seg_off, dtype=np.int64),
              hit_ray=np.array(hit_ray, dtype=np.int64), hit_group=np.array([h[0] for h in hit_rows], dtype=np.int32),
              hit_points=np.array([h[1] for h in hit_rows]).reshape(-1, 3), hit_directions=np.array([h[2] for h in hit_rows]).reshape(-1, 3),
              hit_powers=np.array([h[3] for h in hit_rows]), hit_is_entering=np.array([h[4] for h in hit_rows], dtype=np.uint8))


def main():
  ps, ray_mod, og = reference_modules()
  oracle = Oracle()
  out = {}
  for name, n in cases.FIXTURE_CASES.items():
    sim = cases.fixture_case(name)
    rec = sim.source_records[0]
    src = sim.source_args(0)
    drawn = oracle.sample_mc(src, cases.SEED, 0, n)                      # (theta | r, phi) of rays 0..n-1: the sampler is pinned separately
    light = Obj(Name=rec['name'], Label=rec['label'], FocalLength=rec['FocalLength'], Wavelength=float(rec['Wavelength']),
                MaxRayLengthScale=float(rec['MaxRayLengthScale']), MaxIntersectionsScale=float(rec['MaxIntersectionsScale']))
    proxy = ps.PointSourceProxy.__new__(ps.PointSourceProxy)
    gpM = Matrix(rec['gpM'])
    proxy._getCoordinateTransformMatricesWithoutLinks = lambda obj: (gpM, gpM.inverse(), gpM, gpM.inverse())
    rays = [proxy._makeRay(light, float(t), float(p)) for t, p in zip(drawn['first'], drawn['phi'])]   # the reference's own _makeRay
    out[name+'/origins'] = np.array([tuple(r.initPoint) for r in rays])
    out[name+'/directions'] = np.array([tuple(r.initDirection) for r in rays])
    out[name+'/first'], out[name+'/phi'] = drawn['first'], drawn['phi']
    out[name+'/wavelength'] = np.array(float(rec['Wavelength']))
    cfg = sim.cfg(record_all_hits=True, wavelength=float(rec['Wavelength']))
    objs = make_objects(og, sim.scene)
    ignored = [sim.scene.group_names.index(g) if isinstance(g, str) else int(g) for g in rec['ignored']]
    res = run_case(ray_mod, oracle, sim.scene, objs, cfg, rays, light, ignored)
    out.update({f'{name}/{k}': v for k, v in res.items()})
    print(name, n, 'rays', len(res['seg_power']), 'segments', len(res['hit_powers']), 'interactions', flush=True)
  for name, (build, wavelengths) in cases.SYNTHETIC_CASES.items():
    scene, o, d, settings = build()
    objs = make_objects(og, scene)
    for wl in wavelengths:
      key = name if len(wavelengths) == 1 else f'{name}@{wl:g}'
      cfg = cases.synthetic_cfg(settings, record_all_hits=True, wavelength=wl)
      light = Obj(Name='Source', Label='Source', Wavelength=wl, MaxRayLengthScale=1.0, MaxIntersectionsScale=1.0)
      rays = [ray_mod.Ray(light, Vector(a), Vector(b), wavelength=wl) for a, b in zip(o, d)]
      res = run_case(ray_mod, oracle, scene, objs, cfg, rays, light)
      out[key+'/origins'], out[key+'/directions'], out[key+'/wavelength'] = o, d, np.array(wl)
      out.update({f'{key}/{k}': v for k, v in res.items()})
      print(key, len(o), 'rays', len(res['seg_power']), 'segments', len(res['hit_powers']), 'interactions', flush=True)
  path = os.path.join(HERE, 'traceray_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
