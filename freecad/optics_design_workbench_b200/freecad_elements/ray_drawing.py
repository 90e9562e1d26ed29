'''
GUI ray drawing from device segment buffers (reference freecad_elements/generic_source.py:96-138).

The reference draws while it traces: every segment a ray yields becomes a Part line, appended to a `RaySegment`
Part::Feature of the light source (lines in the source's LOCAL coordinates, gpMi * p, :105; compound of lines per feature,
:127-134; LineWidth / LineColor from the source's view object, :121-124).  Here the rays are traced on the GPU first (every
intersection recorded), turned into per-ray polylines (GenericSourceProxy._ray_dicts) and then handed to a drawing back end:

  FreeCADBackend   the reference's calls, for use inside FreeCAD (Part.makeLine / makeCompound, document.addObject,
                   obj.ElementList); one RaySegment feature per ray.  The per-object colour changes of the reference
                   (optical_group.py:327-339, view-proxy colouring) are not applied: a ray keeps the source's colour.
  RecordingBackend keeps the line lists (tests, headless use: e.g. to export rays to another viewer)
'''
import numpy as np


class RecordingBackend:
  'collects what would be drawn: one entry per ray = array (n_segments, 2, 3) of line end points in the source\'s local frame'
  def __init__(self):
    self.cleared = 0
    self.rays = []

  def clear(self, source):
    self.cleared += 1
    self.rays = []

  def add_ray(self, source, lines):
    self.rays.append(np.asarray(lines, dtype=np.float64))


class FreeCADBackend:
  'the reference\'s drawing calls (generic_source.py:56,105-134); needs FreeCAD'
  def __init__(self, document=None):
    import FreeCAD, Part                              # raises ImportError outside FreeCAD: the caller reports it
    self.App, self.Part = FreeCAD, Part
    self.document = document or FreeCAD.ActiveDocument

  def clear(self, source):
    for o in list(getattr(source, 'ElementList', [])):
      if o.Name.startswith('RaySegment'):
        self.document.removeObject(o.Name)

  def add_ray(self, source, lines):
    V = self.App.Vector
    segs = [self.Part.makeLine(V(*a), V(*b)) for a, b in lines]
    if not segs:
      return
    o = self.document.addObject('Part::Feature', 'RaySegment')
    o.Visibility = False
    vo, so = getattr(o, 'ViewObject', None), getattr(source, 'ViewObject', None)
    if vo is not None and so is not None:
      vo.ShowInTree = False
      vo.LineWidth = so.LineWidth
      vo.LineColor = so.ShapeMaterial.DiffuseColor
    o.Shape = self.Part.makeCompound(segs)
    source.ElementList = source.ElementList + [o]
    o.Visibility = True


def default_backend():
  try:
    return FreeCADBackend()
  except ImportError as e:
    raise RuntimeError('draw=True creates FreeCAD Part objects and FreeCAD is not importable here; pass drawBackend= '
                       '(e.g. ray_drawing.RecordingBackend()) to receive the line lists instead') from e


def draw_rays(source_object, ray_dicts, gpM, backend):
  '''
  ray_dicts: per-ray polylines in WORLD coordinates (dict(points (M+1, 3), ...), GenericSourceProxy._ray_dicts);
  gpM: the light source's global placement — lines are drawn in its local frame (gpMi * p, generic_source.py:105).
  '''
  gpMi = np.linalg.inv(np.asarray(gpM, dtype=np.float64).reshape(4, 4))
  backend.clear(source_object)
  for r in ray_dicts:
    pts = np.asarray(r['points'], dtype=np.float64)
    local = pts @ gpMi[:3, :3].T + gpMi[:3, 3]
    backend.add_ray(source_object, np.stack([local[:-1], local[1:]], axis=1))
