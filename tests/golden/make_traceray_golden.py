'''
Generates tests/golden/traceray_golden.npz by running the reference's OWN Python for the hot path
  PointSourceProxy._makeRay            (freecad_elements/point_source.py:411-460)
  Ray.traceRay / getNormal / mirror / snellsLaw / lineGrating   (freecad_elements/ray.py:36-281,455-539)
  Ray.findNearestIntersection          (freecad_elements/ray.py:290-452: candidate shells by enlarged-box distance, the LINE
                                        test of the boxes, face candidates, the three acceptance rules, the maxRayLength
                                        shrink, the minDist + 2 tol filter, the "not the current medium" preference)
  find.relevantOpticalObjects          (freecad_elements/find.py:79-104: ignore list, sequential filter) with
  SimulationSettingsProxy.getTracingSequence (freecad_elements/simulation_settings.py:158-196)
  raytracing_cache.cached*             (simulation/raytracing_cache.py:43-114)
  OpticalGroupProxy.onRayHit / applyStochasticRayCorrections   (freecad_elements/optical_group.py:206-209,279-323;
                                        its (theta, phi) draws come from the engine's Philox stream: numpy's process-seeded
                                        RNG of the reference is not reproducible — the rotation formula is the reference's)
imported unmodified from /root/reference in the build container.  FreeCAD is absent, so
  * FreeCAD.Vector / Rotation / Matrix are the stand-ins of tests/freecad_stub.py,
  * the document is a list of plain objects (optical groups, one settings object) behind simulation.simulatingDocument(),
  * the OpenCASCADE primitives the loop calls — Part.makeLine, Curve.intersect(Surface), Part.Vertex.distToShape(edge | face),
    BoundBox.isInside / closestPoint / enlarge / intersect, Surface.parameter, Face.normalAt — are the stand-ins of
    tests/occ_stub.py (numpy; polynomial root finding, not the closed forms of the oracle or the kernels; no use of oracle/).
NO method of the reference's Ray is overridden: the golden holds what the reference's code decides, given OCC-primitive
answers only.  What stays unpinned is OpenCASCADE itself (the primitive answers on real BRep shapes).
Run here only (/root/reference does not exist on the GPU box):  python tests/golden/make_traceray_golden.py
'''
import os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

import freecad_stub
freecad_stub.install()
import occ_stub
occ_stub.install(sys.modules['Part'])
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


class Document:
  'what simulation.simulatingDocument() returns: find._allObjects walks .Objects (find.py:24-56)'
  def __init__(self, objects):
    self.Objects = list(objects)


def make_settings(ss, cfg, objs, sequence):
  'the active OpticalSimulationSettings object (simulation_settings.py:20-77), read by find.activeSimulationSettings'
  st = Obj(Name='OpticalSimulationSettings', Label='OpticalSimulationSettings', TypeId='Part::FeaturePython', Active=True,
           Proxy=ss.SimulationSettingsProxy.__new__(ss.SimulationSettingsProxy),
           MaxRayLength=cfg.cfg.max_ray_length, MaxIntersections=cfg.cfg.max_intersections,
           DistanceTolerance=repr(float(cfg.cfg.dist_tol)), SequentialMode=bool(sequence))
  st.isDerivedFrom = lambda type_id: False
  st.addProperty = lambda ptype, name, group, doc: setattr(st, name, [])       # a new property list starts empty
  for i, step in enumerate(sequence or []):
    setattr(st, f'SequentialModeElements_{i:02d}', [objs[g] for g in step])
  return st


def run_case(ray_mod, oracle, scene, objs, cfg, rays, light, ignored=(), sequence=None, frames=None, shape_scene=None):
  '''
  rays: list of reference Ray objects.  Returns per-ray segment lists and the hits handed to the store.
  The reference's own findNearestIntersection / getNormal run on occ_stub shapes built from the scene's face table
  (frames / shape_scene: the shapes live in the groups' own coordinates and are reached through gpM, pM).
  '''
  sa = _abi.SceneArgs(scene)
  ss = sys.modules['odw_ref.freecad_elements.simulation_settings']
  identity = Matrix()
  for o in objs:
    o.TypeId = 'App::LinkGroupPython'
    o.isDerivedFrom = lambda type_id: False
    o.Shape = occ_stub.group_shape(shape_scene if shape_scene is not None else scene, o.group_index)
    T, S = (frames or {}).get(o.group_index, (np.eye(4), np.eye(4)))
    gpM, pM = Matrix(T), Matrix(S)
    o.Proxy._getCoordinateTransformMatrices = lambda obj, m=(gpM, gpM.inverse(), pM, pM.inverse()): [m]
  settings = make_settings(ss, cfg, objs, sequence)
  sys.modules['odw_ref.simulation'].simulatingDocument = lambda: Document(objs + [settings])
  sys.modules['odw_ref.simulation.raytracing_cache'].cacheClear()
  light.IgnoredOpticalElements = [objs[g] for g in ignored]

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

  found = []

  class TracedRay(ray_mod.Ray):
    'observes the reference method (bounce counter for the Philox draws, which face was chosen); changes nothing'
    def findNearestIntersection(self, *a, **k):
      state['bounce'] += 1
      hit = ray_mod.Ray.findNearestIntersection(self, *a, **k)
      if hit is not None:
        found.append(hit[1][2].index)
      return hit

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
  assert len(found) == len(hit_rows)
  return dict(seg_p1=np.array(seg_p1).reshape(-1, 3), seg_p2=np.array(seg_p2).reshape(-1, 3), seg_power=np.array(seg_power),
              seg_medium=np.array(seg_medium, dtype=np.int32), seg_offsets=np.array(seg_off, dtype=np.int64),
              hit_ray=np.array(hit_ray, dtype=np.int64), hit_group=np.array([h[0] for h in hit_rows], dtype=np.int32),
              hit_face_id=scene.faces['face_id'][np.array(found, dtype=np.int64)].astype(np.int32),
              hit_face_row=np.array(found, dtype=np.int32),
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
    sc = sim.scene
    sequence = [[int(g) for g in sc.seq_groups[a:b]] for a, b in zip(sc.seq_offsets[:-1], sc.seq_offsets[1:])] if cfg.cfg.sequential else None
    res = run_case(ray_mod, oracle, sim.scene, objs, cfg, rays, light, ignored, sequence=sequence)
    out.update({f'{name}/{k}': v for k, v in res.items()})
    print(name, n, 'rays', len(res['seg_power']), 'segments', len(res['hit_powers']), 'interactions', flush=True)
  for name, (build, wavelengths) in cases.SYNTHETIC_CASES.items():
    scene, o, d, settings = build()
    _, extras = cases.split_settings(settings)
    objs = make_objects(og, scene)
    for wl in wavelengths:
      key = name if len(wavelengths) == 1 else f'{name}@{wl:g}'
      cfg = cases.synthetic_cfg(settings, record_all_hits=True, wavelength=wl)
      light = Obj(Name='Source', Label='Source', Wavelength=wl, MaxRayLengthScale=1.0, MaxIntersectionsScale=1.0)
      rays = [ray_mod.Ray(light, Vector(a), Vector(b), wavelength=wl) for a, b in zip(o, d)]
      res = run_case(ray_mod, oracle, scene, objs, cfg, rays, light, ignored=extras['ignored'] or (), sequence=extras['sequence'],
                     frames=extras['frames'], shape_scene=extras['shape_scene'])
      out[key+'/origins'], out[key+'/directions'], out[key+'/wavelength'] = o, d, np.array(wl)
      out.update({f'{key}/{k}': v for k, v in res.items()})
      print(key, len(o), 'rays', len(res['seg_power']), 'segments', len(res['hit_powers']), 'interactions', flush=True)
  path = os.path.join(HERE, 'traceray_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
