'''
Turns a saved FreeCAD project (or a stored scene fixture) into everything one simulation needs:
flat scene, per-source sampler tables + descriptors, and the POD trace configuration.  This is the
one-shot export that replaces the reference's per-segment document walks
(reference freecad_elements/find.py:79-141, simulation/raytracing_cache.py:43-114) — settings are
read once here instead of once per segment (reference freecad_elements/ray.py:46,286).
'''

import json

import numpy as np

from .. import _abi
from ..distributions import point_source_tables, surface_source_tables
from ..scene_export import fcstd
from ..scene_export.scene import Scene, SRC_POINT_SPHERICAL, SRC_POINT_COLLIMATED, SRC_SURFACE


class PreparedSimulation:
  '''
  scene            scene_export.scene.Scene
  settings         dict (MaxRayLength, MaxIntersections, DistanceTolerance, SequentialMode, …)
  source_records   list of dicts (PowerDensity, domains, resolutions, gpM, …)
  '''
  def __init__(self, scene, settings, source_records, info=None):
    self.scene, self.settings, self.source_records = scene, settings, source_records
    self.info = info or {}
    self._source_args = {}

  # -- constructors
  @classmethod
  def from_fcstd(cls, path):
    doc = fcstd.FCStdDocument(path)
    scene, info = fcstd.build_scene(doc)
    info['program_version'] = doc.program_version
    return cls(scene, fcstd.settings_dict(doc), fcstd.source_records(doc), info)

  @classmethod
  def from_fixture(cls, path):
    'scene fixture written by save_fixture (tests/golden/scenes/*.npz): the exported scene, not the FCStd'
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z['meta']))
    scatters = []
    for i in range(int(meta.get('n_scatters', 0))):
      from ..distributions import SamplerTables
      d = meta['scatter_domains'][i]
      scatters.append(SamplerTables(z[f'scatter_phi_{i}'], z[f'scatter_first_{i}'], d[0], d[1], 'theta'))
    from ..scene_export.scene import GROUP_DTYPE
    groups = z['groups']
    if groups.dtype != GROUP_DTYPE and groups.dtype.itemsize == GROUP_DTYPE.itemsize:
      groups = groups.view(GROUP_DTYPE)                      # fixtures written before a field of the same size was renamed
    scene = Scene(z['faces'], z['segs'], z['shells'], groups, meta['group_names'], meta['group_labels'],
                  z['seq_offsets'], z['seq_groups'], scatters=scatters,
                  group_scatter=z['group_scatter'] if 'group_scatter' in z.files else None)
    records = meta['source_records']
    for i, r in enumerate(records):
      r['gpM'] = np.array(r['gpM'], dtype=np.float64)
      if f'emit_faces_{i}' in z.files:
        from ..freecad_elements.surface_source import EmittingFaces
        r['emit'] = EmittingFaces(z[f'emit_faces_{i}'], z[f'emit_segs_{i}'])
    settings = meta['settings']
    for k in ('EndAfterRays', 'EndAfterHits', 'EndAfterIterations'):
      if settings.get(k) is None:
        settings[k] = np.inf
    return cls(scene, settings, records, dict(fixture=str(path)))

  def save_fixture(self, path):
    def clean(v):
      if isinstance(v, np.ndarray):
        return v.tolist()
      if isinstance(v, (np.floating, float)):
        return None if not np.isfinite(v) else float(v)
      if isinstance(v, (np.integer,)):
        return int(v)
      return v
    meta = dict(group_names=self.scene.group_names, group_labels=self.scene.group_labels,
                settings={k: clean(v) for k, v in self.settings.items()},
                source_records=[{k: clean(v) for k, v in r.items() if k != 'emit'} for r in self.source_records])
    extra = {}
    if self.scene.scatters:
      meta['n_scatters'] = len(self.scene.scatters)
      meta['scatter_domains'] = [[list(t.first_domain), list(t.phi_domain)] for t in self.scene.scatters]
      extra['group_scatter'] = self.scene.group_scatter
      for i, t in enumerate(self.scene.scatters):
        extra[f'scatter_phi_{i}'], extra[f'scatter_first_{i}'] = t.phi_cdf, t.first_cdf
    for i, r in enumerate(self.source_records):
      if r.get('emit') is not None:
        extra[f'emit_faces_{i}'], extra[f'emit_segs_{i}'] = r['emit'].faces, r['emit'].segs
    np.savez_compressed(path, faces=self.scene.faces, segs=self.scene.segs, shells=self.scene.shells,
                        groups=self.scene.groups, seq_offsets=self.scene.seq_offsets,
                        seq_groups=self.scene.seq_groups, meta=np.array(json.dumps(meta)), **extra)

  # -- engine inputs
  def cfg(self, **overrides):
    'odw_trace_cfg from the active settings (ray.py:46-73,283-288)'
    kw = dict(max_ray_length=self.settings['MaxRayLength'],
              dist_tol=max(self.settings['DistanceTolerance'], 1e-6),
              power_tol=1e-6,
              max_intersections=int(self.settings['MaxIntersections']),
              sequential=self.settings['SequentialMode'])
    kw.update(overrides)
    return _abi.CfgArgs(**kw)

  def source_args(self, index=0):
    'SourceArgs (tables + odw_source_desc) of light source `index`; tables are built once'
    if index not in self._source_args:
      rec = self.source_records[index]
      if rec['proxy'] == 'SurfaceSourceProxy':
        emit = rec.get('emit')
        if emit is None or not len(emit.faces):
          raise NotImplementedError(f"surface source {rec['name']}: no emitting faces the engine can sample "
                                    f"({rec.get('emit_error') or 'ActiveSurfaces empty'})")
        tables = surface_source_tables(rec)
        self._source_args[index] = _abi.SourceArgs(
          tables, kind=SRC_SURFACE, source_id=rec['source_id'], gpM=np.eye(4), wavelength=float(rec['Wavelength']),
          ignored=rec['ignored'], max_ray_length_scale=float(rec['MaxRayLengthScale']),
          max_intersections_scale=float(rec['MaxIntersectionsScale']),
          emit_faces=emit.faces, emit_segs=emit.segs, emit_cdf=emit.cdf,
          dist_tol=max(float(self.settings['DistanceTolerance']), 1e-9))        # surface_source.py:113-119
        return self._source_args[index]
      if rec['proxy'] != 'PointSourceProxy':
        raise NotImplementedError(f"source kind {rec['proxy']} has no device sampler yet")
      tables = point_source_tables(rec)
      f = float(rec['FocalLength'])
      kind = SRC_POINT_SPHERICAL if np.isfinite(f) else SRC_POINT_COLLIMATED
      self._source_args[index] = _abi.SourceArgs(
        tables, kind=kind, source_id=rec['source_id'], gpM=rec['gpM'],
        focal_length=f if np.isfinite(f) else 0.0, wavelength=float(rec['Wavelength']),
        ignored=rec['ignored'], max_ray_length_scale=float(rec['MaxRayLengthScale']),
        max_intersections_scale=float(rec['MaxIntersectionsScale']))
    return self._source_args[index]


def prepare(path):
  path = str(path)
  if path.endswith('.npz'):
    return PreparedSimulation.from_fixture(path)
  return PreparedSimulation.from_fcstd(path)
