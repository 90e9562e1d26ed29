'''
ctypes mirror of include/odw.h (struct layouts only, no library loading).  Shared by the product
binding (`_lib.py`) and by the test-side oracle loader so both describe scenes identically.
'''

import ctypes as C
import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint64_p = C.POINTER(C.c_uint64)
c_uint8_p = C.POINTER(C.c_uint8)

ODW_OK, ODW_EINVAL, ODW_ENODEVICE, ODW_ECUDA, ODW_ENOMEM, ODW_EOVERFLOW, ODW_EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6


class SceneDesc(C.Structure):
  _fields_ = [('n_faces', C.c_int32), ('n_segs', C.c_int32), ('n_shells', C.c_int32), ('n_groups', C.c_int32),
              ('n_seq_steps', C.c_int32), ('n_seq_entries', C.c_int32),
              ('faces', C.c_void_p), ('segs', C.c_void_p), ('shells', C.c_void_p), ('groups', C.c_void_p),
              ('seq_offsets', C.c_void_p), ('seq_groups', C.c_void_p),
              ('n_scatters', C.c_int32), ('pad0', C.c_int32), ('scatters', C.c_void_p), ('group_scatter', C.c_void_p)]


class Scatter(C.Structure):
  _fields_ = [('n_first', C.c_int32), ('n_phi', C.c_int32), ('n_rows', C.c_int32), ('n_tables', C.c_int32),
              ('first_lo', C.c_double), ('first_hi', C.c_double), ('phi_lo', C.c_double), ('phi_hi', C.c_double),
              ('phi_cdf', C.c_void_p), ('first_cdf', C.c_void_p)]


class SourceDesc(C.Structure):
  _fields_ = [('kind', C.c_int32), ('source_id', C.c_int32), ('n_first', C.c_int32), ('n_phi', C.c_int32),
              ('n_rows', C.c_int32), ('n_ignored', C.c_int32),
              ('first_lo', C.c_double), ('first_hi', C.c_double), ('phi_lo', C.c_double), ('phi_hi', C.c_double),
              ('focal_length', C.c_double), ('wavelength', C.c_double),
              ('max_ray_length_scale', C.c_double), ('max_intersections_scale', C.c_double),
              ('gpM', C.c_double*16),
              ('phi_cdf', C.c_void_p), ('first_cdf', C.c_void_p), ('ignored_groups', C.c_void_p),
              ('n_emit', C.c_int32), ('n_emit_segs', C.c_int32),
              ('emit_faces', C.c_void_p), ('emit_segs', C.c_void_p), ('emit_cdf', C.c_void_p), ('dist_tol', C.c_double)]


class Binning(C.Structure):
  _fields_ = [('group', C.c_int32), ('nu', C.c_int32), ('nv', C.c_int32), ('weighted', C.c_int32),
              ('origin', C.c_double*3), ('uaxis', C.c_double*3), ('vaxis', C.c_double*3),
              ('u_lo', C.c_double), ('u_hi', C.c_double), ('v_lo', C.c_double), ('v_hi', C.c_double)]


class TraceCfg(C.Structure):
  _fields_ = [('max_ray_length', C.c_double), ('dist_tol', C.c_double), ('power_tol', C.c_double),
              ('max_intersections', C.c_int32), ('sequential', C.c_int32), ('record_all_hits', C.c_int32),
              ('store_hits', C.c_int32), ('n_binnings', C.c_int32), ('bounces_per_wave', C.c_int32),
              ('hit_capacity', C.c_uint64), ('binnings', C.c_void_p), ('wavelength', C.c_double), ('scatter_seed', C.c_uint64)]


class Counts(C.Structure):
  _fields_ = [(k, C.c_uint64) for k in ('rays', 'segments', 'hits', 'hits_dropped', 'escaped',
                                         'depth_terminated', 'waves', 'sm_clock_khz')]

  def as_dict(self):
    return {k: int(getattr(self, k)) for k, _ in self._fields_ if k != 'sm_clock_khz'}


class HitsView(C.Structure):
  _fields_ = [('capacity', C.c_uint64), ('points', C.c_void_p), ('directions', C.c_void_p), ('powers', C.c_void_p),
              ('is_entering', C.c_void_p), ('ray_index', C.c_void_p), ('group', C.c_void_p), ('bounce', C.c_void_p),
              ('face_id', C.c_void_p), ('medium', C.c_void_p)]


def _ptr(a):
  return a.ctypes.data if a is not None and a.size else None


class SceneArgs:
  'keeps the numpy arrays alive next to the ctypes struct that points into them'
  def __init__(self, scene):
    self.scene = scene
    d = SceneDesc()
    d.n_faces, d.n_segs = len(scene.faces), len(scene.segs)
    d.n_shells, d.n_groups = len(scene.shells), len(scene.groups)
    d.n_seq_steps, d.n_seq_entries = scene.n_seq_steps, len(scene.seq_groups)
    d.faces, d.segs = _ptr(scene.faces), _ptr(scene.segs)
    d.shells, d.groups = _ptr(scene.shells), _ptr(scene.groups)
    d.seq_offsets, d.seq_groups = _ptr(scene.seq_offsets), _ptr(scene.seq_groups)
    tables = list(getattr(scene, 'scatters', []) or [])
    if tables:
      self.scatters = (Scatter*len(tables))()
      for sct, t in zip(self.scatters, tables):
        sct.n_first, sct.n_phi, sct.n_rows = t.first_cdf.shape[-1], t.phi_cdf.shape[-1], t.first_cdf.shape[-2]
        sct.n_tables = int(getattr(t, 'n_tables', 1))
        sct.first_lo, sct.first_hi = t.first_domain
        sct.phi_lo, sct.phi_hi = t.phi_domain
        sct.phi_cdf, sct.first_cdf = _ptr(t.phi_cdf), _ptr(t.first_cdf)
      self.group_scatter = np.ascontiguousarray(scene.group_scatter, dtype=np.int32).reshape(len(scene.groups), 2)
      d.n_scatters, d.scatters, d.group_scatter = len(tables), C.addressof(self.scatters), _ptr(self.group_scatter)
    self.desc = d


class SourceArgs:
  def __init__(self, tables, *, kind, source_id, gpM, focal_length=0.0, wavelength=500.0, ignored=(),
               max_ray_length_scale=1.0, max_intersections_scale=1.0, emit_faces=None, emit_segs=None,
               emit_cdf=None, dist_tol=1e-6):
    self.phi_cdf = np.ascontiguousarray(tables.phi_cdf, dtype=np.float64)
    self.first_cdf = np.ascontiguousarray(tables.first_cdf, dtype=np.float64)
    self.ignored = np.ascontiguousarray(list(ignored), dtype=np.int32)
    d = SourceDesc()
    d.kind, d.source_id = kind, source_id
    d.n_first, d.n_phi = self.first_cdf.shape[-1], self.phi_cdf.shape[0]
    d.n_rows = 1 if self.first_cdf.ndim == 1 else self.first_cdf.shape[0]
    d.n_ignored = len(self.ignored)
    d.first_lo, d.first_hi = tables.first_domain
    d.phi_lo, d.phi_hi = tables.phi_domain
    d.focal_length, d.wavelength = focal_length, wavelength
    d.max_ray_length_scale, d.max_intersections_scale = max_ray_length_scale, max_intersections_scale
    m = np.asarray(gpM, dtype=np.float64).reshape(16)
    for i in range(16):
      d.gpM[i] = m[i]
    d.phi_cdf, d.first_cdf = _ptr(self.phi_cdf), _ptr(self.first_cdf)
    d.ignored_groups = _ptr(self.ignored)
    if emit_faces is not None:                      # surface source (ODW_SRC_SURFACE)
      self.emit_faces = np.ascontiguousarray(emit_faces)
      self.emit_segs = np.ascontiguousarray(emit_segs)
      self.emit_cdf = np.ascontiguousarray(emit_cdf, dtype=np.float64)
      d.n_emit, d.n_emit_segs = len(self.emit_faces), len(self.emit_segs)
      d.emit_faces, d.emit_segs, d.emit_cdf = _ptr(self.emit_faces), _ptr(self.emit_segs), _ptr(self.emit_cdf)
      d.dist_tol = float(dist_tol)
    self.desc = d
    self.tables = tables


class CfgArgs:
  def __init__(self, *, max_ray_length=1000.0, dist_tol=1e-6, power_tol=1e-6, max_intersections=100,
               sequential=False, record_all_hits=False, store_hits=True, binnings=(), bounces_per_wave=0,
               hit_capacity=0, scatter_seed=0, wavelength=0.0):
    self.binnings = (Binning*max(1, len(binnings)))()
    for i, b in enumerate(binnings):
      bb = self.binnings[i]
      bb.group, bb.nu, bb.nv, bb.weighted = b['group'], b['nu'], b['nv'], int(b.get('weighted', 0))
      for j in range(3):
        bb.origin[j], bb.uaxis[j], bb.vaxis[j] = b['origin'][j], b['uaxis'][j], b['vaxis'][j]
      bb.u_lo, bb.u_hi, bb.v_lo, bb.v_hi = b['u_range'][0], b['u_range'][1], b['v_range'][0], b['v_range'][1]
    self.binning_specs = list(binnings)
    c = TraceCfg()
    c.max_ray_length = max_ray_length
    c.dist_tol = max(float(dist_tol), 1e-6)          # ray.py:288
    c.power_tol = power_tol
    c.max_intersections = int(max_intersections)
    c.sequential = int(bool(sequential))
    c.record_all_hits = int(bool(record_all_hits))
    c.store_hits = int(bool(store_hits))
    c.n_binnings = len(binnings)
    c.bounces_per_wave = int(bounces_per_wave)
    c.hit_capacity = int(hit_capacity)
    c.binnings = C.addressof(self.binnings) if len(binnings) else None
    c.scatter_seed = int(scatter_seed)
    c.wavelength = float(wavelength)
    self.cfg = c


class HitArrays:
  'host arrays for a hit list + the odw_hits_view pointing at them'
  def __init__(self, capacity):
    capacity = int(capacity)
    self.points = np.empty((capacity, 3), dtype=np.float64)
    self.directions = np.empty((capacity, 3), dtype=np.float64)
    self.powers = np.empty(capacity, dtype=np.float64)
    self.is_entering = np.empty(capacity, dtype=np.uint8)
    self.ray_index = np.empty(capacity, dtype=np.uint64)
    self.group = np.empty(capacity, dtype=np.int32)
    self.bounce = np.empty(capacity, dtype=np.int32)
    self.face_id = np.empty(capacity, dtype=np.int32)
    self.medium = np.empty(capacity, dtype=np.int32)
    v = HitsView()
    v.capacity = capacity
    for k in ('points', 'directions', 'powers', 'is_entering', 'ray_index', 'group', 'bounce', 'face_id', 'medium'):
      setattr(v, k, getattr(self, k).ctypes.data)
    self.view = v
    self.n = 0

  def trimmed(self, n, sort=True):
    n = int(n)
    self.n = n
    out = {k: getattr(self, k)[:n] for k in ('points', 'directions', 'powers', 'is_entering', 'ray_index',
                                             'group', 'bounce', 'face_id', 'medium')}
    if sort and n:
      order = np.lexsort((out['bounce'], out['ray_index']))
      out = {k: v[order] for k, v in out.items()}
    return out
