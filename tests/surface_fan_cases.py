'''
Faces and face sets of the surface-source fan-mode golden (tests/golden/make_surface_fan_golden.py writes it,
tests/test_surface_fan.py reads it).
'''
import numpy as np

from freecad.optics_design_workbench_b200.freecad_elements.surface_source import emitting_faces_from_instances
from freecad.optics_design_workbench_b200.scene_export import primitives as prim


def _only(face_instances, index, transform=None):
  return lambda: emitting_faces_from_instances([([face_instances()[index]], np.eye(4) if transform is None else transform)])


def _all(parts):
  return lambda: emitting_faces_from_instances([(make(), tr) for make, tr in parts])


TILT = prim.translation(3, -2, 7) @ prim.rotation((1, 2, 0.5), 0.7)

# name -> (emitting face builder, distance tolerance, requested grid point counts)
GRID_CASES = {
  'rectangle': (_only(lambda: prim.rectangle(10, 4), 0, TILT), 1e-2, (1, 4, 9, 25, 100)),
  'disc': (_only(lambda: prim.disc(6.0), 0), 1e-2, (1, 4, 9, 30, 200)),
  'sphere': (_only(lambda: prim.sphere(5.0), 0, TILT), 1e-2, (1, 4, 9, 50, 300)),
  'cylinder_side': (_only(lambda: prim.cylinder(2.0, 9.0), 0), 1e-2, (4, 9, 40)),
  'torus': (_only(lambda: prim.torus(10.0, 2.0), 0), 1e-3, (9, 60)),
  'cone_side': (_only(lambda: prim.cone(4.0, 1.0, 6.0), 0), 1e-2, (9, 45)),
  'lens_cap': (_only(lambda: prim.plano_convex_lens(5.0, 2.0), 0), 1e-2, (4, 20, 100)),
}

# name -> (emitting face set builder, distance tolerance, FanModeRayCount)
SOURCE_CASES = {
  'box_100': (_all([(lambda: prim.box(10, 10, 1), prim.translation(-5, -5, 38))]), 1e-2, 100),
  'box_and_sphere_60': (_all([(lambda: prim.box(4, 6, 2), TILT), (lambda: prim.sphere(3.0), prim.translation(0, 0, -20))]), 1e-2, 60),
  'lens_30': (_all([(lambda: prim.plano_convex_lens(5.0, 2.0), np.eye(4))]), 1e-2, 30),
  'many_faces_10': (_all([(lambda: prim.box(1, 2, 3), prim.translation(4*i, 0, 0)) for i in range(4)]), 1e-2, 10),
}
