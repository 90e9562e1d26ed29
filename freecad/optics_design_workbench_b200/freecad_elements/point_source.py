'''
Host side of the point light source: which rays one simulation iteration consists of.

Mirrors PointSourceProxy of the reference (reference freecad_elements/point_source.py):
  _makeRay        :411-460   (theta|r, phi) -> global origin + unit direction      -> make_rays (vectorised)
  _generateRays   :463-682   'fans'  : deterministic fans, generated HERE on the host and traced through
                                       odw_trace_rays (explicit ray list)
                             'true'  : Monte-Carlo draws — NOT generated here: the kernel draws them itself from
                                       Philox counters (odw_trace_mc); this module only says how many
`obj` is anything with the reference's property names as attributes or keys (a FreeCAD document object
inside FreeCAD, a source record of scene_export.fcstd.source_records here).
'''

import numpy as np
import sympy as sy

from ..distributions import sampler_tables as st


def _get(obj, key, default=None):
  if isinstance(obj, dict):
    return obj.get(key, default)
  return getattr(obj, key, default)


class RayBatch:
  '''
  Explicit ray list in global coordinates (what the reference holds as a list of ray.Ray objects):
  origins [n,3], directions [n,3] (unit), powers [n], wavelength, and the per-ray metadata the
  reference attaches in _makeRay / _generateRays (initPhi, initTheta, initRadius, fanIndex, rayIndex,
  totalFanCount, totalRaysInFan).
  '''
  def __init__(self, origins, directions, powers, wavelength, metadata):
    self.origins = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
    self.directions = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    self.powers = np.ascontiguousarray(powers, dtype=np.float64)
    self.wavelength = float(wavelength)
    self.metadata = {k: np.asarray(v) for k, v in metadata.items()}

  def __len__(self):
    return len(self.origins)


def make_rays(obj, gpM, theta_or_radius, phi, power=1.0, metadata=None):
  '''
  PointSourceProxy._makeRay (point_source.py:411-460), vectorised.
  finite f:  d = Rz(phi) Rx(theta) z = (sin t sin p, -sin t cos p, cos t),  o = f (z - d)
  f = inf :  d = z,  o = r (cos p, -sin p, 0)      [orthoAxis x opticalAxis = (0,-1,0)]
  then both through the source's global placement, direction renormalised.
  '''
  first = np.atleast_1d(np.asarray(theta_or_radius, dtype=np.float64))
  phi = np.atleast_1d(np.asarray(phi, dtype=np.float64))
  f = float(_get(obj, 'FocalLength', '0'))
  n = len(first)
  if np.isfinite(f):
    theta = first
    radius = np.tan(theta)*f
    ld = np.stack([np.sin(theta)*np.sin(phi), -np.sin(theta)*np.cos(phi), np.cos(theta)], axis=1)
    lo = (np.array([0.0, 0.0, 1.0])[None, :]-ld)*f
  else:
    radius = first
    theta = np.full(n, np.nan)
    ld = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    lo = np.stack([radius*np.cos(phi), -radius*np.sin(phi), np.zeros(n)], axis=1)
  M = np.asarray(gpM, dtype=np.float64).reshape(4, 4)
  p1 = lo @ M[:3, :3].T + M[:3, 3]
  p2 = (lo + ld/np.linalg.norm(ld, axis=1)[:, None]) @ M[:3, :3].T + M[:3, 3]
  d = p2-p1
  d /= np.linalg.norm(d, axis=1)[:, None]
  md = dict(initPhi=phi, initTheta=theta, initRadius=radius)
  md.update(metadata or {})
  return RayBatch(p1, d, np.full(n, float(power)), float(_get(obj, 'Wavelength', 500.0)), md)


def _fan_sides(obj, fan_phi, l1, l2, fan_mode, rays_per_fan, phi_dom, first_var):
  '''
  One fan: ([(index, value, phi)] in the order the reference yields the rays, l1, l2), or (None, l1, l2) if the fan
  is skipped (point_source.py:512-656).  l1, l2 come back because the reference clips them to the FanModePowerSpan
  INSIDE its loop over fans and never resets them (:546-549), so every fan starts from the limits the previous
  fan left behind and the span is applied again on top — a quirk we reproduce (pinned by tests/golden/fan_golden.npz).
  '''
  density = str(_get(obj, 'PowerDensity'))
  f = float(_get(obj, 'FocalLength', '0'))
  phiL1, phiL2 = phi_dom
  cands = [p for p in np.arange(fan_phi-30*np.pi, fan_phi+31*np.pi, np.pi) if phiL1-1e-9 <= p <= phiL2+1e-9]
  if not cands:
    return None, l1, l2
  phiA = cands[int(np.argmin(np.abs(fan_phi-np.array(cands))))]
  cands = [p for p in np.arange(phiA+np.pi-30*np.pi, phiA+np.pi+31*np.pi, 2*np.pi) if phiL1-1e-9 <= p <= phiL2+1e-9]
  phiB = np.nan if not cands else cands[int(np.argmin(np.abs(phiA+np.pi-np.array(cands))))]

  span = float(_get(obj, 'FanModePowerSpan', 0.9))
  if 0 < span < 1:                                                    # :534-549
    power_vs = sy.lambdify(first_var, sy.sympify(density).subs('theta', 'abs(theta)')
                           .subs('phi', f'Piecewise( ( ({phiA}), ({first_var})>0 ), ( ({phiB}),  True     ) )'))
    limit = max(abs(l1), abs(l2))
    grid = np.linspace(-limit, limit, int(1e5))
    cum = np.cumsum(power_vs(grid)*np.ones_like(grid))
    cum = cum/max(cum)
    _l1 = grid[int(np.argmin(np.abs(cum-(1-span)/2)))]
    _l2 = grid[int(np.argmin(np.abs(cum-(1-(1-span)/2))))]
    maxL = max(abs(_l1), abs(_l2))
    if abs(l1) > maxL:
      l1 = np.sign(l1)*maxL
    if abs(l2) > maxL:
      l2 = np.sign(l2)*maxL

  res = float(_get(obj, 'ThetaResolutionNumericMode' if first_var == 'theta' else 'RadiusResolutionNumericMode', '1e5'))

  def grid_for(expr_string, domain, phi_value, n):
    expr, _ = st.point_source_density(expr_string, f, scalar=True)
    return st.find_grid(expr, first_var, domain, res, n, constants=dict(phi=phi_value))

  if fan_mode == 'gapped':
    side1 = grid_for(density, (l1, l2), phiA, rays_per_fan//2)
    side2 = grid_for(density, (l1, l2), phiB, rays_per_fan//2)
  elif fan_mode == 'stitched':
    limit = max(abs(l1), abs(l2))
    e = sy.sympify(density).subs('theta', 'abs(theta)').subs('r', 'abs(r)')
    if np.isfinite(phiB):
      e = e.subs('phi', f'Piecewise( ( ({phiA}), ({first_var})>0 ), ( ({phiB}),  True     ) )')
      dom = (-limit, limit)
    else:
      dom = (0, limit)
    side1 = grid_for(str(e), dom, phiA, rays_per_fan)
    side2 = []
  else:                                                               # theta-sign-change
    side1 = grid_for(density, (l1, l2), phiA, rays_per_fan)
    side2 = []

  if len(side2) > 0:                                                  # :623-629
    side1 = sorted(side1, key=abs)
    side2 = sorted(side2, key=abs)
    idx1 = list(1+np.arange(len(side1)))
    idx2 = list(-(1+np.arange(len(side2))))
  else:                                                               # :634-638
    side1 = np.array(sorted(side1))
    i0 = int(np.argmin(np.abs(side1)))
    idx1 = list(np.arange(len(side1))-i0)
    idx2 = []
  packed = (list(zip(idx1, side1, [phiA]*len(side1))) + list(zip(idx2, side2, [phiB]*len(side2))))
  packed = sorted(packed, key=lambda e: abs(e[0])-.1)                 # stable, like the reference's sorted()
  return packed, l1, l2


def generate_fan_rays(obj, gpM, max_fan_count=np.inf, max_rays_per_fan=np.inf):
  '''
  PointSourceProxy._generateRays(mode='fans') (point_source.py:474-656) -> RayBatch in the reference's order.
  '''
  rays_per_fan = min(_get(obj, 'RaysPerFan', 20), max_rays_per_fan)
  total_fans = int(min(_get(obj, 'Fans', 2), max_fan_count))
  f = float(_get(obj, 'FocalLength', '0'))
  if np.isfinite(f):
    l1, l2 = st.parse_domain(_get(obj, 'ThetaDomain', '0, pi/4'), (0, np.pi/4))
    first_var = 'theta'
  else:
    l1, l2 = st.parse_domain(_get(obj, 'RadiusDomain', '0, 10'), (0, 10))
    first_var = 'r'
  phi_dom = st.parse_domain(_get(obj, 'PhiDomain', '0, 2*pi'), (0, 2*np.pi))
  if (l1 > 0 and l2 > 0) or (l1 < 0 and l2 < 0):
    fan_mode = 'gapped'
    rays_per_fan = max(4, int(np.ceil(rays_per_fan/2)*2))             # even, at least 4 (quirk Q5)
  elif l1 == 0 or l2 == 0:
    fan_mode = 'stitched'
  elif l1 < 0 and l2 > 0:
    fan_mode = 'theta-sign-change'
  else:
    raise ValueError(f'{l1=}, {l2=}')
  phi0 = float(sy.sympify(_get(obj, 'FanPhi0', '0')).evalf())
  first, phis, fan_index, ray_index, n_in_fan = [], [], [], [], []
  for fi, fan_phi in enumerate(phi0 + np.linspace(0, np.pi, total_fans+1)[:-1]):
    packed, l1, l2 = _fan_sides(obj, fan_phi, l1, l2, fan_mode, rays_per_fan, phi_dom, first_var)
    if packed is None:
      continue
    for ri, v, p in packed:
      first.append(v); phis.append(p); fan_index.append(fi); ray_index.append(int(ri)); n_in_fan.append(len(packed))
  md = dict(fanIndex=np.array(fan_index, dtype=np.int64), rayIndex=np.array(ray_index, dtype=np.int64),
            totalFanCount=np.full(len(first), total_fans, dtype=np.int64),
            totalRaysInFan=np.array(n_in_fan, dtype=np.int64))
  return make_rays(obj, gpM, np.array(first, dtype=np.float64), np.array(phis, dtype=np.float64), metadata=md)


def rays_per_iteration(obj, settings):
  "'true'/'pseudo' mode: RaysPerIteration * RaysPerIterationScale (point_source.py:661-665)"
  n = 100 if settings is None else settings.get('RaysPerIteration', 100)
  return int(n*float(_get(obj, 'RaysPerIterationScale', 1.0)))
