'''
TEST DOUBLE: an object with the surface of engine.Engine whose tracing is done by the CPU oracle.
It lets the host-side logic (runSimulationIteration, the hit writer, the sharded simulation loop) run in
the `not gpu` suite.  The product never imports this; product code gets the CUDA engine.
'''
import numpy as np

from oracle import Oracle


class _Result:
  def __init__(self, r):
    self._r = r
    self.overflow = False
    self.kernel_ms = 0.0

  def __enter__(self):
    return self

  def __exit__(self, *a):
    pass

  def close(self):
    pass

  @property
  def counts(self):
    return self._r['counts']

  def hits(self, sort=True, into=None):
    return self._r['hits']

  def histogram(self, index=0):
    return self._r['histograms'][index]

  def ray_summary(self):
    return {k: self._r[k] for k in ('n_segments', 'final_points', 'final_powers', 'final_media')}


class _Scene:
  def __init__(self, orc, scene):
    self.orc, self.scene = orc, scene

  def trace_mc(self, source, cfg, seed, first_ray, n_rays):
    return _Result(self.orc.trace_mc(self.scene, source.args, cfg, seed, first_ray, n_rays,
                                     hit_capacity=max(16, int(cfg.cfg.hit_capacity) or 4*n_rays), threads=0))

  def trace_mc_host(self, source, cfg, seed, first_ray, n_rays, view):
    'same contract as DeviceScene.trace_mc_host: rows go into the arrays behind the odw_hits_view'
    import ctypes as C
    r = self.orc.trace_mc(self.scene, source.args, cfg, seed, first_ray, n_rays, hit_capacity=int(view.capacity), threads=0, sort=False)
    h, got = r['hits'], len(r['hits']['powers'])
    for name, width in (('points', 3), ('directions', 3), ('powers', 1), ('is_entering', 1), ('ray_index', 1), ('group', 1),
                        ('bounce', 1), ('face_id', 1), ('medium', 1)):
      ptr = getattr(view, name)
      if ptr:
        src = np.ascontiguousarray(h[name])
        C.memmove(ptr, src.ctypes.data, src.nbytes)
    return r['counts'], got

  def trace_rays(self, cfg, origins, directions, powers=None, ignored=()):
    return _Result(self.orc.trace_rays(self.scene, cfg, origins, directions, powers, ignored=ignored, threads=0))

  def close(self):
    pass


class _Source:
  def __init__(self, orc, args):
    self.orc, self.args = orc, args

  def sample(self, seed, first_ray, n):
    return self.orc.sample_mc(self.args, seed, first_ray, n)

  def close(self):
    pass


class OracleEngine:
  def __init__(self):
    self.orc = Oracle()

  def scene(self, scene):
    return _Scene(self.orc, scene)

  def source(self, source_args):
    return _Source(self.orc, source_args)

  def close(self):
    pass
