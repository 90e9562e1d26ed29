'''
Headless FCStd importer: reads a saved FreeCAD project (zip of Document.xml + *.brp + binary
property files) WITHOUT FreeCAD and produces the flat scene + light-source + settings
descriptions the engine consumes.

Reference behaviour restated here (reference tree freecad/optics_design_workbench/…):
  freecad_elements/find.py:59-141           which objects are light sources / optical groups / settings
  freecad_elements/common.py:36-125         global placement of an object through nested Parts, groups and links
  freecad_elements/ray.py:332-345           face world transform  gpM * pMi * Shape  (Shape already holds pM)
  freecad_elements/simulation_settings.py:158-196   sequential-mode element lists
Everything FreeCAD would hand the reference as `obj.Shape` is rebuilt from the stored BRep of the
leaf features; links, link arrays (Draft Array with PlacementList) and groups are expanded here.
'''

import base64
import io
import json
import struct
import zipfile
import xml.etree.ElementTree as ET

import numpy as np

from . import brep
from .scene import SceneBuilder, OPTICAL_TYPES


def quat_placement_matrix(px, py, pz, q0, q1, q2, q3):
  'FreeCAD Placement (position + quaternion x,y,z,w) -> 4x4 matrix'
  x, y, z, w = q0, q1, q2, q3
  n = np.sqrt(x*x+y*y+z*z+w*w)
  if n == 0:
    x, y, z, w = 0., 0., 0., 1.
  else:
    x, y, z, w = x/n, y/n, z/n, w/n
  m = np.eye(4)
  m[0, 0] = 1-2*(y*y+z*z); m[0, 1] = 2*(x*y-z*w);   m[0, 2] = 2*(x*z+y*w)
  m[1, 0] = 2*(x*y+z*w);   m[1, 1] = 1-2*(x*x+z*z); m[1, 2] = 2*(y*z-x*w)
  m[2, 0] = 2*(x*z-y*w);   m[2, 1] = 2*(y*z+x*w);   m[2, 2] = 1-2*(x*x+y*y)
  m[:3, 3] = (px, py, pz)
  return m


class DocObject:
  def __init__(self, name, type_id):
    self.Name = name
    self.TypeId = type_id
    self.props = {}
    self.proxy_module = None
    self.proxy_class = None
    self.proxy_state = None

  def get(self, key, default=None):
    return self.props.get(key, default)

  def __repr__(self):
    return f'<{self.TypeId} {self.Name}>'

  @property
  def Label(self):
    return self.props.get('Label', self.Name)

  @property
  def placement(self):
    return self.props.get('Placement', np.eye(4))


class FCStdDocument:
  '''Parsed FCStd archive: objects in document order with decoded properties.'''

  def __init__(self, path):
    self.path = str(path)
    self._zip = zipfile.ZipFile(self.path)
    self._brep_cache = {}
    root = ET.fromstring(self._zip.read('Document.xml'))
    self.program_version = root.attrib.get('ProgramVersion', '?')
    self.objects = {}
    self.order = []
    for o in root.find('Objects').findall('Object'):
      obj = DocObject(o.attrib['name'], o.attrib['type'])
      self.objects[obj.Name] = obj
      self.order.append(obj.Name)
    for od in root.find('ObjectData').findall('Object'):
      obj = self.objects[od.attrib['name']]
      props = od.find('Properties')
      for p in (props if props is not None else []):
        self._read_property(obj, p)

  # -- property decoding
  def _read_property(self, obj, p):
    name, typ = p.attrib.get('name'), p.attrib.get('type')
    kids = list(p)
    if not kids:
      return
    k = kids[0]
    tag = k.tag
    val = None
    if tag == 'PropertyPlacement':
      a = k.attrib
      val = quat_placement_matrix(*(float(a[x]) for x in ('Px', 'Py', 'Pz', 'Q0', 'Q1', 'Q2', 'Q3')))
    elif tag == 'Float':
      val = float(k.attrib['value'])
    elif tag == 'Integer':
      val = int(k.attrib['value'])
      if typ == 'App::PropertyEnumeration':
        enum = p.find('CustomEnumList')
        if enum is not None:
          names = [e.attrib['value'] for e in enum.findall('Enum')]
          if 0 <= val < len(names):
            val = names[val]
    elif tag == 'Bool':
      val = k.attrib['value'].lower() == 'true'
    elif tag == 'String':
      val = k.attrib['value']
    elif tag == 'LinkList':
      val = [l.attrib['value'] for l in k.findall('Link')]
    elif tag == 'Link':
      val = k.attrib.get('value') or None
    elif tag == 'XLink':
      val = k.attrib.get('name') or None
      if k.attrib.get('file'):
        val = None        # cross-document links need the in-FreeCAD exporter
    elif tag == 'LinkSubList':
      val = [(l.attrib.get('obj'), l.attrib.get('sub')) for l in k.findall('Link')]
    elif tag == 'PropertyVector':
      a = k.attrib
      val = np.array([float(a['valueX']), float(a['valueY']), float(a['valueZ'])])
    elif tag == 'Part':
      val = k.attrib.get('file')
    elif tag == 'PlacementList':
      val = self._read_placement_list(k.attrib.get('file'))
    elif tag == 'VectorList':
      val = self._read_vector_list(k.attrib.get('file'))
    elif tag == 'Python':
      obj.proxy_module = k.attrib.get('module')
      obj.proxy_class = k.attrib.get('class')
      try:
        raw = k.attrib.get('value', '')
        if k.attrib.get('encoded') == 'yes':
          raw = base64.b64decode(raw).decode()
        obj.proxy_state = json.loads(raw) if raw else None
      except Exception:
        obj.proxy_state = None
      return
    else:
      return
    obj.props[name] = val

  def _read_placement_list(self, member):
    if not member or member not in self._zip.namelist():
      return []
    d = self._zip.read(member)
    if len(d) < 4:
      return []
    n = struct.unpack('<I', d[:4])[0]
    out = []
    for i in range(n):
      v = struct.unpack('<7d', d[4+56*i:4+56*(i+1)])
      out.append(quat_placement_matrix(*v))
    return out

  def _read_vector_list(self, member):
    if not member or member not in self._zip.namelist():
      return np.zeros((0, 3))
    d = self._zip.read(member)
    if len(d) < 4:
      return np.zeros((0, 3))
    n = struct.unpack('<I', d[:4])[0]
    if len(d) >= 4+24*n:
      return np.frombuffer(d, dtype='<f8', count=3*n, offset=4).reshape(n, 3).copy()
    return np.frombuffer(d, dtype='<f4', count=3*n, offset=4).reshape(n, 3).astype(float)

  def brep_faces(self, member):
    'FaceInstance list of a stored shape (cached); [] when the member is empty/missing'
    if member not in self._brep_cache:
      faces = []
      if member and member in self._zip.namelist():
        text = self._zip.read(member).decode('ascii', errors='replace')
        if text.strip():
          faces = brep.read_brep(text).faces()
      self._brep_cache[member] = faces
    return self._brep_cache[member]

  # -- object classes (find.py:59-141)
  def _is_proxy(self, obj, classes):
    return obj.proxy_class in classes

  def light_sources(self):
    return [self.objects[n] for n in self.order
            if self.objects[n].TypeId == 'App::LinkGroupPython'
            and self.objects[n].proxy_class in ('PointSourceProxy', 'SurfaceSourceProxy', 'ReplaySourceProxy')]

  def optical_groups(self):
    return [self.objects[n] for n in self.order
            if self.objects[n].TypeId == 'App::LinkGroupPython'
            and self.objects[n].proxy_class == 'OpticalGroupProxy']

  def active_settings(self):
    allS = [self.objects[n] for n in self.order
            if self.objects[n].TypeId == 'Part::FeaturePython'
            and self.objects[n].proxy_class == 'SimulationSettingsProxy']
    active = [s for s in allS if s.get('Active', False)]
    if len(active) > 1:
      raise ValueError('only one simulation settings object may be active: '
                       + ', '.join(s.Name for s in active))
    if active:
      return active[0]
    return allS[0] if allS else None

  # -- structure
  def children(self, obj):
    'objects an object claims as geometry children'
    t = obj.TypeId
    if t in ('App::Part', 'App::DocumentObjectGroup', 'PartDesign::Body') or t.startswith('App::DocumentObjectGroup'):
      return list(obj.get('Group', []) or [])
    if t.startswith('App::LinkGroup'):
      return list(obj.get('ElementList', []) or [])
    return []

  def parents(self, obj):
    '''
    Containers claiming obj.  A plain App::DocumentObjectGroup inside an App::Part lists its members a
    second time (the Part's Group holds them too); such a group adds no placement and no new instance,
    so it only counts as a parent when no geometric container claims the object.
    '''
    ps = [self.objects[n] for n in self.order if obj.Name in self.children(self.objects[n])]
    geo = [p for p in ps if not p.TypeId.startswith('App::DocumentObjectGroup')]
    return geo if geo else ps

  def links_to(self, obj):
    return [self.objects[n] for n in self.order
            if self.objects[n].TypeId.startswith('App::Link') and not self.objects[n].TypeId.startswith('App::LinkGroup')
            and self.objects[n].get('LinkedObject') == obj.Name]

  def global_placements(self, obj, ignore_links=False, _depth=0):
    '''
    All 4x4 global placement matrices under which obj exists in the document
    (restates allPlacementsAndPaths, common.py:36-109): through every chain of containing
    Parts / groups, plus once more for every in-document App::Link pointing at obj or at one of its
    containers.  Result entries: (matrix, dotted path).
    '''
    if _depth > 100:
      raise RuntimeError('placement recursion too deep')
    out = []
    parents = self.parents(obj)
    own = obj.placement
    if not parents:
      out.append((own, obj.Name))
    for par in parents:
      for pm, ppath in self.global_placements(par, ignore_links, _depth+1):
        out.append((pm @ own, ppath+'.'+obj.Name))
    if not ignore_links:
      for link in self.links_to(obj):
        lt = bool(link.get('LinkTransform', False))
        elems = link.get('PlacementList', []) if int(link.get('ElementCount', 0) or 0) > 0 else [np.eye(4)]
        for lm, lpath in self.global_placements(link, ignore_links, _depth+1):
          for em in elems:
            out.append((lm @ em @ (own if lt else np.eye(4)), lpath+'.'+obj.Name))
    if _depth == 0:
      out = sorted(out, key=lambda e: e[1])
    return out

  # -- shapes
  def shape_instances(self, obj, with_placement=True, _depth=0):
    '''
    What FreeCAD returns as obj.Shape, as a list of (FaceInstance list, 4x4 matrix) pairs: the
    matrix is applied on top of the stored BRep (which already carries the feature's own
    Placement as its root location).
    '''
    if _depth > 100:
      raise RuntimeError('shape recursion too deep')
    t = obj.TypeId
    P = obj.placement
    Pi = np.linalg.inv(P)
    out = []
    is_link = t.startswith('App::Link') and not t.startswith('App::LinkGroup')
    is_draft_array = (obj.proxy_module or '').startswith('draftobjects.') and obj.get('Base') and \
        obj.get('PlacementList') is not None and not obj.get('ExpandArray', False)
    if is_link or is_draft_array:
      target = self.objects.get(obj.get('LinkedObject') if is_link else obj.get('Base'))
      if target is None:
        return []
      lt = bool(obj.get('LinkTransform', False))
      scale = obj.get('Scale', 1.0)
      if scale is not None and abs(float(scale)-1) > 1e-12:
        raise NotImplementedError(f'{obj.Name}: link scale {scale} is not supported by the headless importer')
      base = self.shape_instances(target, with_placement=lt, _depth=_depth+1)
      n = int(obj.get('ElementCount', 0) or 0) if is_link else len(obj.get('PlacementList') or [])
      elems = (obj.get('PlacementList') or [])[:n] if n > 0 else [np.eye(4)]
      for em in elems:
        for faces, m in base:
          out.append((faces, (P if with_placement else np.eye(4)) @ em @ m))
      return out
    kids = self.children(obj)
    if kids and not obj.get('Shape'):
      for name in kids:
        child = self.objects.get(name)
        if child is None or child.Name.startswith('RaySegment'):
          continue
        for faces, m in self.shape_instances(child, True, _depth+1):
          has_p = 'Placement' in obj.props
          out.append((faces, (P if (with_placement and has_p) else np.eye(4)) @ m))
      return out
    member = obj.get('Shape')
    if member:
      faces = self.brep_faces(member)
      if faces:
        out.append((faces, np.eye(4) if with_placement else Pi))
    return out


# ------------------------------------------------------------------------------------------

def _float(s, default):
  try:
    return float(s)
  except Exception:
    return default


def settings_dict(doc):
  '''
  Tunables of the active OpticalSimulationSettings object with the reference's defaults when no
  settings object exists (ray.py:46-73,283-288; simulation_settings.py:20-77).
  '''
  s = doc.active_settings()
  d = dict(MaxRayLength=1000.0, MaxIntersections=100.0, DistanceTolerance=1e-2, RaysPerIteration=100.0,
           SequentialMode=False, EndAfterRays=np.inf, EndAfterHits=np.inf, EndAfterIterations=np.inf,
           store_hit_keys=[], name=None)
  if s is None:
    return d
  d['name'] = s.Name
  d['MaxRayLength'] = float(s.get('MaxRayLength', 1000.0))
  d['MaxIntersections'] = float(s.get('MaxIntersections', 100.0))
  d['DistanceTolerance'] = _float(s.get('DistanceTolerance', '1e-6'), 1e-6)
  d['RaysPerIteration'] = float(s.get('RaysPerIteration', 100.0))
  d['SequentialMode'] = bool(s.get('SequentialMode', False))
  for k in ('EndAfterRays', 'EndAfterHits', 'EndAfterIterations'):
    d[k] = _float(s.get(k, 'inf'), np.inf)
  d['store_hit_keys'] = [k[8:] for k, v in s.props.items() if k.startswith('StoreHit') and v is True]
  return d


def tracing_sequence(doc):
  'non-empty SequentialModeElements_NN lists in ascending order (simulation_settings.py:158-196)'
  s = doc.active_settings()
  if s is None or not s.get('SequentialMode', False):
    return []
  seq = []
  for i in range(100):
    lst = s.get(f'SequentialModeElements_{i:02d}')
    if lst:
      seq.append(list(lst))
  return seq


def build_scene(doc):
  '''
  FCStdDocument -> (Scene, info).  Group order = document order of optical groups (find.py:69-76).
  Every (placement of the group) x (element of the group) x (instance of the element) becomes world
  faces:  world = gpM(group) * element.Shape   — equal to the reference's gpM*pMi*group.Shape because
  group.Shape = pM * compound(element shapes)  (ray.py:332-339).
  '''
  b = SceneBuilder()
  groups = doc.optical_groups()
  index = {}
  for g in groups:
    otype = g.get('OpticalType')
    if otype not in OPTICAL_TYPES:
      otype = (g.proxy_state or {}).get('oldType', 'Vacuum')
    orient = g.get('GratingLinesOrientation')
    gi = b.add_group(
      g.Name, g.Label, otype,
      refractive_index=float(g.get('RefractiveIndex', 2.0)),
      reflectivity=float(g.get('Reflectivity', 1.0)),
      absorption_length=_float(g.get('AbsorptionLength', 'inf'), np.inf),
      record_hits=bool(g.get('RecordHits', False)),
      fresnel=bool(g.get('FresnelReflection', False)),       # additional property of this engine (default off = reference behaviour)
      grating_type=g.get('GratingType', 'Reflection') if g.get('GratingType') in ('Reflection', 'Transmission') else 'Reflection',
      grating_lines_per_mm=float(g.get('GratingLinesPerMillimeter', 1000.0)),
      grating_order=float(g.get('GratingDiffractionOrder', 1.0)),
      grating_orientation=orient if orient is not None else (0, 0, 1),
      scatter_density=(g.get('ReflectedProbabilityDensity', '') if otype == 'Mirror'
                       else g.get('RefractedProbabilityDensity', '') if otype == 'Lens' else ''),
      power_theta_domain=g.get('PowerThetaDomain', '-pi/2, pi/2'), power_phi_domain=g.get('PowerPhiDomain', '0, 2*pi'),
      modify_density=g.get('RayModificationProbabilityDensity', ''),
      modify_theta_domain=g.get('ModifyThetaDomain', '-pi/2, pi/2'), modify_phi_domain=g.get('ModifyPhiDomain', '0, 2*pi'))
    index[g.Name] = gi
    for gpM, _path in doc.global_placements(g):
      for name in g.get('ElementList', []) or []:
        child = doc.objects.get(name)
        if child is None:
          continue
        for faces, m in doc.shape_instances(child):
          b.add_shape(gi, faces, gpM @ m)
  sequence = [[index[n] for n in step if n in index] for step in tracing_sequence(doc)]
  sequence = [s for s in sequence if s]
  scene = b.build(sequence)
  info = dict(skipped=b.skipped, group_index=index)
  return scene, info


def source_records(doc):
  '''
  One dict per light source with the properties the sampler/engine needs (point_source.py:32-70,
  generic_source.py:23-37) and its global placement without links (common.py:270-280).
  '''
  out = []
  group_index = {g.Name: i for i, g in enumerate(doc.optical_groups())}
  for i, s in enumerate(doc.light_sources()):
    gp = doc.global_placements(s, ignore_links=True)
    rec = dict(name=s.Name, label=s.Label, proxy=s.proxy_class, source_id=i, gpM=gp[0][0],
               ignored=[group_index[n] for n in (s.get('IgnoredOpticalElements') or []) if n in group_index])
    if s.proxy_class == 'SurfaceSourceProxy':
      try:
        rec['emit'], rec['emit_error'] = surface_source_faces(doc, s), None
      except Exception as e:          # e.g. B-spline emitters: need the tessellation path (not built yet)
        rec['emit'], rec['emit_error'] = None, f'{type(e).__name__}: {e}' 
    for k, default in (('PowerDensity', 'exp(-theta^2/0.01)'), ('Wavelength', 500.0), ('FocalLength', '0'),
                       ('ThetaDomain', '0, pi/4'), ('PhiDomain', '0, 2*pi'), ('RadiusDomain', '0, 10'),
                       ('ThetaResolutionNumericMode', '1e5'), ('RadiusResolutionNumericMode', '1e5'),
                       ('PhiResolutionNumericMode', '1e2'), ('Fans', 2), ('FanPhi0', '0'), ('RaysPerFan', 20),
                       ('FanModePowerSpan', 0.9), ('RaysPerIterationScale', 1.0), ('MaxIntersectionsScale', 1.0),
                       ('MaxRayLengthScale', 1.0), ('RecordRays', False), ('FanModeRayCount', 100), ('ReplayFromDir', '')):
      rec[k] = s.get(k, default)
    out.append(rec)
  return out


def surface_source_faces(doc, source):
  '''
  Emitting faces of a surface source (surface_source.py:437-458): every (part, [FaceN...]) entry of ActiveSurfaces,
  once per placement of the part; an empty face list = all faces of the part.  World transform of a face =
  gpM * pMi * Shape = (global placement of the part) * (shape without the part's own placement).
  '''
  from ..freecad_elements import surface_source
  wanted = {}
  for part_name, sub in source.get('ActiveSurfaces') or []:
    wanted.setdefault(part_name, [])
    if sub:
      wanted[part_name].append(sub)
  selections = []
  for part_name, subs in wanted.items():
    part = doc.objects.get(part_name)
    if part is None:
      continue
    for gpM, _path in doc.global_placements(part):
      instances = doc.shape_instances(part, with_placement=False)
      all_faces = [(fi, m) for faces, m in instances for fi in faces]       # Shape.Faces order
      if subs:
        picked = [all_faces[int(sname[4:])-1] for sname in subs if sname.startswith('Face') and 0 < int(sname[4:]) <= len(all_faces)]
      else:
        picked = all_faces
      for fi, m in picked:
        selections.append(([fi], gpM @ m))
  return surface_source.emitting_faces_from_instances(selections)


def load_fcstd(path):
  'convenience: path -> (Scene, source records, settings dict, info)'
  doc = FCStdDocument(path)
  scene, info = build_scene(doc)
  info['program_version'] = doc.program_version
  return scene, source_records(doc), settings_dict(doc), info
