'''
CPU oracle loader — TEST INFRASTRUCTURE ONLY (see oracle/odw_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package.  The product package never does.  Pinned against the reference's own Python, executed unmodified under stand-ins
for FreeCAD's Base types and the OpenCASCADE primitives it calls: the whole per-ray path — traceRay,
findNearestIntersection, getNormal, the interaction formulas, find.relevantOpticalObjects, _makeRay
(tests/golden/make_traceray_golden.py) — the sampler and the fan grids (tests/golden/), and the Monte-Carlo chi-square
gate (tests/golden/make_mc_gate_golden.py).  "parity unpinned" only for OpenCASCADE's primitive answers on real BRep
shapes (FreeCAD is absent here and on the GPU box; the reference holds no golden ray vectors) — see the header of
odw_oracle.c.
'''

import ctypes as C
import os
import subprocess

import numpy as np

from freecad.optics_design_workbench_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libodw_oracle.so')


def build(force=False):
  src = os.path.join(_HERE, 'odw_oracle.c')
  hdr = os.path.join(_HERE, '..', 'include', 'odw.h')
  stale = (not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
  if force or stale:
    subprocess.run(['make', '-C', _HERE, '-s', '-B', 'all'], check=True)
  return _SO


class Oracle:
  def __init__(self):
    # a prebuilt library travels to the GPU box; rebuild only when the sources are newer and gcc exists
    try:
      build()
    except Exception:
      if not os.path.exists(_SO):
        raise
    self.lib = C.CDLL(_SO)
    L = self.lib
    L.oracle_philox.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p]
    L.oracle_sample_uniforms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    L.oracle_sample_mc.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64] + [C.c_void_p]*4
    L.oracle_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                    C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.oracle_trace_mc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.oracle_max_threads.restype = C.c_int
    L.oracle_find_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_int32,
                                      C.c_void_p, C.c_int32, C.c_void_p]
    L.oracle_face_normal.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.oracle_scatter_draw.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int32,
                                      C.c_void_p, C.c_void_p]

  def find_nearest(self, scene_args, cfg, start, direction, medium, max_len, seq_index, ignored=()):
    'one Ray.findNearestIntersection: (face index, point) or (-1, None)'
    s, d, P = np.array(start, dtype=np.float64), np.array(direction, dtype=np.float64), np.empty(3)
    ign = np.ascontiguousarray(list(ignored), dtype=np.int32)
    fi = self.lib.oracle_find_nearest(C.addressof(scene_args.desc), C.addressof(cfg.cfg), s.ctypes.data, d.ctypes.data,
                                      int(medium), float(max_len), int(seq_index), ign.ctypes.data if len(ign) else None,
                                      len(ign), P.ctypes.data)
    return (fi, P) if fi >= 0 else (-1, None)

  def face_normal(self, scene_args, face, point):
    'Surface.parameter + normalAt of a face at a point: ((u, v), unit normal)'
    P, uv, n = np.array(point, dtype=np.float64), np.empty(2), np.empty(3)
    rc = self.lib.oracle_face_normal(C.addressof(scene_args.desc), int(face), P.ctypes.data, uv.ctypes.data, n.ctypes.data)
    if rc:
      raise ValueError(f'face {face} out of range')
    return uv, n

  def scatter_draw(self, scene_args, group, which, seed, source_id, ray, bounce):
    '(theta, phi) of the stochastic surface model for one interaction, or None when that density is empty'
    th, ph = C.c_double(0), C.c_double(0)
    if not self.lib.oracle_scatter_draw(C.addressof(scene_args.desc), int(group), int(which), int(seed), int(source_id),
                                        int(ray), int(bounce), C.byref(th), C.byref(ph)):
      return None
    return th.value, ph.value

  def max_threads(self):
    return int(self.lib.oracle_max_threads())

  def philox(self, seed, source_id, ray, purpose=0):
    u = np.empty(2)
    self.lib.oracle_philox(int(seed), int(source_id), int(ray), int(purpose), u.ctypes.data)
    return u

  def sample_uniforms(self, source, u_phi, u_first):
    u_phi = np.ascontiguousarray(u_phi, dtype=np.float64)
    u_first = np.ascontiguousarray(u_first, dtype=np.float64)
    first, phi = np.empty_like(u_phi), np.empty_like(u_phi)
    self.lib.oracle_sample_uniforms(C.addressof(source.desc), u_phi.ctypes.data, u_first.ctypes.data,
                                    len(u_phi), first.ctypes.data, phi.ctypes.data)
    return first, phi

  def sample_mc(self, source, seed, first_ray, n):
    first, phi = np.empty(n), np.empty(n)
    o, d = np.empty((n, 3)), np.empty((n, 3))
    self.lib.oracle_sample_mc(C.addressof(source.desc), int(seed), int(first_ray), int(n),
                              first.ctypes.data, phi.ctypes.data, o.ctypes.data, d.ctypes.data)
    return dict(first=first, phi=phi, origins=o, directions=d)

  @staticmethod
  def _bins(cfg):
    sizes = [b['nu']*b['nv'] for b in cfg.binning_specs]
    return np.zeros(max(1, sum(sizes))), sizes

  @staticmethod
  def _split_bins(bins, cfg, sizes):
    out, off = [], 0
    for b, n in zip(cfg.binning_specs, sizes):
      out.append(bins[off:off+n].reshape(b['nu'], b['nv']).copy())
      off += n
    return out

  def trace_rays(self, scene, cfg, origins, directions, powers=None, wavelength=None, ignored=(),
                 hit_capacity=None, threads=1, ray_index_base=0, sort=True):
    sa = scene if isinstance(scene, _abi.SceneArgs) else _abi.SceneArgs(scene)
    if wavelength is None:
      wavelength = cfg.cfg.wavelength if cfg.cfg.wavelength > 0 else 500.0
    o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
    d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    n = len(o)
    p = None if powers is None else np.ascontiguousarray(powers, dtype=np.float64)
    ign = np.ascontiguousarray(list(ignored), dtype=np.int32)
    cap = int(hit_capacity if hit_capacity is not None else max(16, n*(cfg.cfg.max_intersections if cfg.cfg.record_all_hits else 4)))
    hits = _abi.HitArrays(cap)
    nh = C.c_uint64(0)
    counts = _abi.Counts()
    nseg = np.zeros(n, dtype=np.int32)
    fp, fpow, fmed = np.zeros((n, 3)), np.zeros(n), np.zeros(n, dtype=np.int32)
    bins, sizes = self._bins(cfg)
    rc = self.lib.oracle_trace_rays(C.addressof(sa.desc), C.addressof(cfg.cfg), o.ctypes.data, d.ctypes.data,
                                    None if p is None else p.ctypes.data, float(wavelength),
                                    ign.ctypes.data if len(ign) else None, len(ign), n, int(ray_index_base),
                                    C.addressof(hits.view), C.addressof(nh), C.addressof(counts),
                                    nseg.ctypes.data, fp.ctypes.data, fpow.ctypes.data, fmed.ctypes.data, bins.ctypes.data, int(threads))
    return dict(rc=rc, hits=hits.trimmed(nh.value, sort), counts=counts.as_dict(), n_segments=nseg,
                final_points=fp, final_powers=fpow, final_media=fmed, histograms=self._split_bins(bins, cfg, sizes))

  def trace_mc(self, scene, source, cfg, seed, first_ray, n, hit_capacity=None, threads=1, sort=True):
    sa = scene if isinstance(scene, _abi.SceneArgs) else _abi.SceneArgs(scene)
    cap = int(hit_capacity if hit_capacity is not None else max(16, n*4))
    hits = _abi.HitArrays(cap if cfg.cfg.store_hits else 1)
    nh = C.c_uint64(0)
    counts = _abi.Counts()
    bins, sizes = self._bins(cfg)
    rc = self.lib.oracle_trace_mc(C.addressof(sa.desc), C.addressof(source.desc), C.addressof(cfg.cfg),
                                  int(seed), int(first_ray), int(n), C.addressof(hits.view), C.addressof(nh),
                                  C.addressof(counts), bins.ctypes.data, int(threads))
    return dict(rc=rc, hits=hits.trimmed(nh.value if cfg.cfg.store_hits else 0, sort), counts=counts.as_dict(),
                histograms=self._split_bins(bins, cfg, sizes))
