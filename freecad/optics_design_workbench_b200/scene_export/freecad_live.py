'''
In-FreeCAD scene export: builds the flat scene from LIVE document objects (inside a running FreeCAD, where the
reference's own helpers exist) instead of from the saved .FCStd.

It deliberately asks FreeCAD for very little, so that everything geometric stays on the path the tests cover here:
  * `group.Shape.exportBrepToString()`  — the group's compound as OCC's ASCII BRep, parsed by scene_export/brep.py
    exactly like the `*.brp` members of a saved project;
  * the placement matrices the reference itself computes for the group,
    `freecad_elements.common.allCoordinateTransformMatrices(group)` -> [gpM, gpMi, pM, pMi] per placement
    (reference common.py:112-125), handed in as 4x4 arrays.
World transform of a face = gpM * pMi * Shape, as in Ray.findNearestIntersection (reference ray.py:332-345).
Group order = order of the given objects (reference find.py:69-76 yields document order).

FreeCAD is not available in the build container: the unit test drives this with stand-in objects whose
exportBrepToString() returns the stored BRep of a benchmark project.
'''

import numpy as np

from . import brep
from .scene import SceneBuilder, OPTICAL_TYPES


def matrix_to_array(m):
  'FreeCAD.Matrix (attributes A11..A44) or anything array-like -> 4x4 numpy array'
  if hasattr(m, 'A11'):
    return np.array([[getattr(m, f'A{r}{c}') for c in range(1, 5)] for r in range(1, 5)], dtype=np.float64)
  return np.asarray(m, dtype=np.float64).reshape(4, 4)


def _prop(obj, name, default):
  try:
    v = getattr(obj, name)
  except Exception:
    return default
  return default if v is None else v


def build_scene(optical_groups, placements_of, sequence=None):
  '''
  optical_groups   OpticalGroup document objects (find.opticalObjects())
  placements_of    callable(group) -> list of [gpM, gpMi, pM, pMi] (common.allCoordinateTransformMatrices)
  sequence         list of lists of group Names (SimulationSettingsProxy.getTracingSequence, simulation_settings.py:158-196)
                   or None when SequentialMode is off
  Returns (Scene, info).
  '''
  b = SceneBuilder()
  index = {}
  for g in optical_groups:
    otype = _prop(g, 'OpticalType', 'Vacuum')
    if otype not in OPTICAL_TYPES:
      otype = 'Vacuum'
    orient = _prop(g, 'GratingLinesOrientation', (0, 0, 1))
    orient = tuple(getattr(orient, k) for k in 'xyz') if hasattr(orient, 'x') else tuple(orient)
    gi = b.add_group(
      g.Name, g.Label, otype,
      refractive_index=float(_prop(g, 'RefractiveIndex', 2.0)), reflectivity=float(_prop(g, 'Reflectivity', 1.0)),
      absorption_length=float(_prop(g, 'AbsorptionLength', 'inf')), record_hits=bool(_prop(g, 'RecordHits', False)),
      fresnel=bool(_prop(g, 'FresnelReflection', False)),
      grating_type=_prop(g, 'GratingType', 'Reflection'), grating_lines_per_mm=float(_prop(g, 'GratingLinesPerMillimeter', 1000.0)),
      grating_order=float(_prop(g, 'GratingDiffractionOrder', 1.0)), grating_orientation=orient,
      scatter_density=(_prop(g, 'ReflectedProbabilityDensity', '') if otype == 'Mirror'
                       else _prop(g, 'RefractedProbabilityDensity', '') if otype == 'Lens' else ''),
      power_theta_domain=_prop(g, 'PowerThetaDomain', '-pi/2, pi/2'), power_phi_domain=_prop(g, 'PowerPhiDomain', '0, 2*pi'),
      modify_density=_prop(g, 'RayModificationProbabilityDensity', ''),
      modify_theta_domain=_prop(g, 'ModifyThetaDomain', '-pi/2, pi/2'), modify_phi_domain=_prop(g, 'ModifyPhiDomain', '0, 2*pi'))
    index[g.Name] = gi
    faces = brep.read_brep(g.Shape.exportBrepToString()).faces()
    for mats in placements_of(g):
      gpM, _gpMi, _pM, pMi = (matrix_to_array(m) for m in mats)
      b.add_shape(gi, faces, gpM @ pMi)
  seq = None
  if sequence:
    seq = [[index[n] for n in step if n in index] for step in sequence]
    seq = [s for s in seq if s]
  return b.build(seq), dict(skipped=b.skipped, tessellated=b.tessellated, group_index=index)
