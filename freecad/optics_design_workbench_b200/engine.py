'''
ctypes binding of libodw_b200.so (include/odw.h) — the thin host layer between the Python plugin
surface and the sm_100a kernels.  There is no CPU fallback here: a missing library or a missing
CUDA device raises.

Replaces, on the reference side, the body of GenericSourceProxy.runSimulationIteration
(reference freecad_elements/generic_source.py:51-146): instead of generating Ray objects and
calling Ray.traceRay for each, one call traces a whole range of rays on the GPU.
'''

import ctypes as C
import os

import numpy as np

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('ODW_LIB') or os.path.join(_HERE, 'libodw_b200.so')   # ODW_LIB: developer override (kernel variants)

EXPORTS = ['odw_abi_version', 'odw_last_error', 'odw_engine_create', 'odw_engine_destroy', 'odw_engine_device_name', 'odw_engine_stream',
           'odw_host_alloc', 'odw_host_free',
           'odw_scene_create', 'odw_scene_destroy', 'odw_source_create', 'odw_source_destroy',
           'odw_trace_mc', 'odw_trace_mc_host', 'odw_sample_mc', 'odw_trace_rays', 'odw_result_counts', 'odw_result_hits',
           'odw_result_histogram', 'odw_result_histogram_device', 'odw_result_ray_summary', 'odw_result_ray_media',
           'odw_result_kernel_ms', 'odw_result_destroy']

_lib = None


class EngineError(RuntimeError):
  def __init__(self, code, message):
    super().__init__(f'odw error {code}: {message}')
    self.code = code


def build_library(force=False):
  'compile csrc/*.cu for sm_100a in-tree (nvcc cross-compiles without a GPU)'
  import subprocess
  srcdir = os.path.join(_HERE, 'csrc')
  deps = [os.path.join(srcdir, f) for f in ('odw_kernels.cu', 'odw_wavefront.cu', 'odw_api.cu', 'odw_device.cuh', 'odw_trace.cuh')]
  deps.append(os.path.join(_HERE, '..', '..', 'include', 'odw.h'))
  stale = not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(d) for d in deps)
  if force or stale:
    subprocess.run(['make', '-C', srcdir, '-B', 'all'], check=True)
  return LIB_PATH


def load_library():
  'dlopen libodw_b200.so and declare the prototypes; raises if the library is missing'
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise EngineError(-2, f'{LIB_PATH} is missing: build it with __graft_entry__.build() '
                          f'(make -C {os.path.join(_HERE, "csrc")}); there is no CPU fallback')
  L = C.CDLL(LIB_PATH)
  vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int32
  L.odw_abi_version.restype = C.c_int
  L.odw_last_error.restype = C.c_char_p
  L.odw_engine_create.argtypes = [C.c_int, C.POINTER(vp)]
  L.odw_engine_destroy.argtypes = [vp]; L.odw_engine_destroy.restype = None
  L.odw_engine_device_name.argtypes = [vp, C.c_char_p, C.c_int]
  L.odw_engine_stream.argtypes = [vp, C.POINTER(vp)]
  L.odw_host_alloc.argtypes = [vp, u64, C.POINTER(vp)]
  L.odw_host_free.argtypes = [vp, vp]; L.odw_host_free.restype = None
  L.odw_scene_create.argtypes = [vp, vp, C.POINTER(vp)]
  L.odw_scene_destroy.argtypes = [vp]; L.odw_scene_destroy.restype = None
  L.odw_source_create.argtypes = [vp, vp, C.POINTER(vp)]
  L.odw_source_destroy.argtypes = [vp]; L.odw_source_destroy.restype = None
  L.odw_trace_mc.argtypes = [vp, vp, vp, u64, u64, u64, C.POINTER(vp)]
  L.odw_trace_mc_host.argtypes = [vp, vp, vp, u64, u64, u64, vp, C.POINTER(u64), vp]
  L.odw_sample_mc.argtypes = [vp, u64, u64, u64, vp, vp, vp, vp]
  L.odw_trace_rays.argtypes = [vp, vp, vp, vp, vp, vp, i32, u64, C.POINTER(vp)]
  L.odw_result_counts.argtypes = [vp, vp]
  L.odw_result_hits.argtypes = [vp, vp, C.c_int, C.POINTER(u64)]
  L.odw_result_histogram.argtypes = [vp, i32, vp]
  L.odw_result_histogram_device.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(u64)]
  L.odw_result_ray_summary.argtypes = [vp, vp, vp, vp]
  L.odw_result_ray_media.argtypes = [vp, vp]
  L.odw_result_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
  L.odw_result_destroy.argtypes = [vp]; L.odw_result_destroy.restype = None
  _lib = L
  return L


def device_count():
  'number of CUDA devices the library can open (0 without a driver / device)'
  L = load_library()
  n = 0
  while n < 64:
    h = C.c_void_p()
    rc = L.odw_engine_create(n, C.byref(h))
    if rc != 0:
      break
    L.odw_engine_destroy(h)
    n += 1
  return n


def _check(rc, allow=()):
  if rc != 0 and rc not in allow:
    raise EngineError(rc, load_library().odw_last_error().decode(errors='replace'))
  return rc


class TraceResult:
  'owns an odw_result handle'
  def __init__(self, handle, cfg, overflow=False):
    self._h = handle
    self._cfg = cfg
    self.overflow = overflow

  def close(self):
    if getattr(self, '_h', None):
      load_library().odw_result_destroy(self._h)
      self._h = None

  __del__ = close

  def __enter__(self):
    return self

  def __exit__(self, *a):
    self.close()

  @property
  def counts(self):
    c = _abi.Counts()
    _check(load_library().odw_result_counts(self._h, C.addressof(c)))
    return c.as_dict()

  @property
  def sm_clock_mhz(self):
    'effective SM clock the trace kernel saw (clock64 / globaltimer of CTA 0)'
    c = _abi.Counts()
    _check(load_library().odw_result_counts(self._h, C.addressof(c)))
    return c.sm_clock_khz/1e3

  @property
  def kernel_ms(self):
    ms = C.c_double()
    _check(load_library().odw_result_kernel_ms(self._h, C.byref(ms)))
    return ms.value

  def hits(self, sort=True, into=None):
    '''
    hit list as dict of numpy arrays (points, directions, powers, is_entering, ray_index, group, bounce,
    face_id); `into` = a preallocated _abi.HitArrays (e.g. over pinned memory) to copy into.
    '''
    n = min(self.counts['hits'], int(self._cfg.cfg.hit_capacity) or self.counts['hits'])
    arrays = into if into is not None else _abi.HitArrays(max(1, n))
    got = C.c_uint64(0)
    _check(load_library().odw_result_hits(self._h, C.addressof(arrays.view), 1 if sort else 0, C.byref(got)))
    return arrays.trimmed(got.value, sort=False)

  def histogram(self, index=0):
    spec = self._cfg.binning_specs[index]
    out = np.zeros((spec['nu'], spec['nv']), dtype=np.float64)
    _check(load_library().odw_result_histogram(self._h, index, out.ctypes.data))
    return out

  def histogram_device(self, index=0):
    'device pointer + bin count of a histogram (for an NCCL all-reduce by the caller)'
    ptr, n = C.c_void_p(), C.c_uint64()
    _check(load_library().odw_result_histogram_device(self._h, index, C.byref(ptr), C.byref(n)))
    return ptr.value, n.value

  def ray_summary(self):
    n = self.counts['rays']
    nseg, fp, fpow = np.zeros(n, dtype=np.int32), np.zeros((n, 3)), np.zeros(n)
    _check(load_library().odw_result_ray_summary(self._h, nseg.ctypes.data, fp.ctypes.data, fpow.ctypes.data))
    med = np.zeros(n, dtype=np.int32)
    _check(load_library().odw_result_ray_media(self._h, med.ctypes.data))
    return dict(n_segments=nseg, final_points=fp, final_powers=fpow, final_media=med)


class DeviceScene:
  def __init__(self, engine, scene):
    self.engine = engine
    self.scene = scene
    self._args = _abi.SceneArgs(scene)
    h = C.c_void_p()
    _check(load_library().odw_scene_create(engine._h, C.addressof(self._args.desc), C.byref(h)))
    self._h = h

  def close(self):
    if getattr(self, '_h', None):
      load_library().odw_scene_destroy(self._h)
      self._h = None

  __del__ = close

  def trace_rays(self, cfg, origins, directions, powers=None, ignored=()):
    o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
    d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    if o.shape != d.shape:
      raise ValueError('origins and directions must have the same shape')
    p = None if powers is None else np.ascontiguousarray(powers, dtype=np.float64)
    ign = np.ascontiguousarray(list(ignored), dtype=np.int32)
    h = C.c_void_p()
    rc = _check(load_library().odw_trace_rays(self._h, C.addressof(cfg.cfg), o.ctypes.data, d.ctypes.data,
                                              None if p is None else p.ctypes.data,
                                              ign.ctypes.data if len(ign) else None, len(ign), len(o), C.byref(h)),
                allow=(_abi.ODW_EOVERFLOW,))
    return TraceResult(h, cfg, overflow=(rc == _abi.ODW_EOVERFLOW))

  def trace_mc_host(self, source, cfg, seed, first_ray, n_rays, hits_view):
    '''
    odw_trace_mc_host: hit rows are written into the host arrays behind `hits_view` (an _abi.HitsView, ideally
    over page-locked memory) while later chunks are still being traced.  Returns (counts dict, rows written).
    '''
    got, counts = C.c_uint64(0), _abi.Counts()
    _check(load_library().odw_trace_mc_host(self._h, source._h, C.addressof(cfg.cfg), int(seed), int(first_ray),
                                            int(n_rays), C.addressof(hits_view), C.byref(got), C.addressof(counts)),
           allow=(_abi.ODW_EOVERFLOW,))
    return counts.as_dict(), got.value

  def trace_mc(self, source, cfg, seed, first_ray, n_rays):
    h = C.c_void_p()
    rc = _check(load_library().odw_trace_mc(self._h, source._h, C.addressof(cfg.cfg), int(seed), int(first_ray),
                                            int(n_rays), C.byref(h)), allow=(_abi.ODW_EOVERFLOW,))
    return TraceResult(h, cfg, overflow=(rc == _abi.ODW_EOVERFLOW))


class DeviceSource:
  def __init__(self, engine, source_args):
    self.engine = engine
    self.args = source_args
    h = C.c_void_p()
    _check(load_library().odw_source_create(engine._h, C.addressof(source_args.desc), C.byref(h)))
    self._h = h

  def close(self):
    if getattr(self, '_h', None):
      load_library().odw_source_destroy(self._h)
      self._h = None

  __del__ = close

  def sample(self, seed, first_ray, n):
    first, phi = np.empty(n), np.empty(n)
    o, d = np.empty((n, 3)), np.empty((n, 3))
    _check(load_library().odw_sample_mc(self._h, int(seed), int(first_ray), int(n), first.ctypes.data,
                                        phi.ctypes.data, o.ctypes.data, d.ctypes.data))
    return dict(first=first, phi=phi, origins=o, directions=d)


class Engine:
  'one engine per GPU (one process per GPU in multi-GPU runs)'
  def __init__(self, device_id=0):
    L = load_library()
    if L.odw_abi_version() != 2:
      raise EngineError(-1, 'ABI version mismatch between engine.py and libodw_b200.so')
    h = C.c_void_p()
    _check(L.odw_engine_create(int(device_id), C.byref(h)))
    self._h = h
    self.device_id = device_id
    self._pinned = []

  def close(self):
    if getattr(self, '_h', None):
      for ptr in self._pinned:
        load_library().odw_host_free(self._h, ptr)
      self._pinned = []
      load_library().odw_engine_destroy(self._h)
      self._h = None

  def device_name(self):
    buf = C.create_string_buffer(256)
    _check(load_library().odw_engine_device_name(self._h, buf, 256))
    return buf.value.decode()

  def stream_handle(self):
    'raw cudaStream_t of the engine (wrap with torch.cuda.ExternalStream to record events on it)'
    st = C.c_void_p()
    _check(load_library().odw_engine_stream(self._h, C.byref(st)))
    return st.value or 0

  def pinned_hit_arrays(self, capacity, columns=('points', 'directions', 'powers', 'is_entering', 'group')):
    '''
    Page-locked host arrays for odw_trace_mc_host: returns (dict of numpy arrays, _abi.HitsView).  Only the listed
    columns are allocated; the others stay NULL and are not copied.  The memory lives as long as the engine.
    '''
    spec = dict(points=(np.float64, 3), directions=(np.float64, 3), powers=(np.float64, 1), is_entering=(np.uint8, 1),
                ray_index=(np.uint64, 1), group=(np.int32, 1), bounce=(np.int32, 1), face_id=(np.int32, 1), medium=(np.int32, 1))
    view = _abi.HitsView()
    view.capacity = int(capacity)
    arrays = {}
    for name in columns:
      dtype, width = spec[name]
      nbytes = int(capacity)*width*np.dtype(dtype).itemsize
      ptr = C.c_void_p()
      _check(load_library().odw_host_alloc(self._h, nbytes, C.byref(ptr)))
      self._pinned.append(ptr)
      buf = (C.c_char*max(nbytes, 1)).from_address(ptr.value)
      a = np.frombuffer(buf, dtype=dtype, count=int(capacity)*width)
      arrays[name] = a.reshape(int(capacity), 3) if width == 3 else a
      setattr(view, name, ptr.value)
    return arrays, view

  def free_pinned(self, view):
    'give the page-locked columns behind `view` (from pinned_hit_arrays) back; the numpy arrays over them die with it'
    for name in ('points', 'directions', 'powers', 'is_entering', 'ray_index', 'group', 'bounce', 'face_id', 'medium'):
      ptr = getattr(view, name)
      ptr = ptr if isinstance(ptr, int) else getattr(ptr, 'value', None)
      if not ptr:
        continue
      for k, have in enumerate(self._pinned):
        if have.value == ptr:
          load_library().odw_host_free(self._h, have)
          del self._pinned[k]
          break
      setattr(view, name, None)

  def scene(self, scene):
    return DeviceScene(self, scene)

  def source(self, source_args):
    return DeviceSource(self, source_args)
