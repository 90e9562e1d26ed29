'''
Host-side construction of the tabulated inverse-CDF sampler and the deterministic fan grid.

Restates the NUMERIC mode of the reference sampler
(reference distributions/random_number_generator.py:337-464 `_generateNumericScalarLambda` /
`_lambdasFromSampled`) and the fan-grid quantile placement
(reference distributions/points_by_density.py:25-38, random_number_generator.py:685-725).  The
tables go to the device once per simulation (odw_source_create); the per-ray draw itself happens
in the CUDA kernel.

Differences to the reference, both deliberate:
  * the analytic (sympy integrate + solve) mode is never used: the tabulated mode is the
    reference's own fallback, gives the same distribution and is the only one that maps to a GPU;
    all benchmark sources report 'numeric' anyway.
  * a density without phi is detected symbolically and stored as ONE conditional row
    (n_rows = 1) instead of 100 identical 100001-entry rows.
'''

import numpy as np
import sympy as sy


def parse_domain(text, default=(0.0, 1.0)):
  'domain string "a, b" with sympy expressions (reference freecad_elements/common.py:293-361, happy path)'
  try:
    vals = [float(sy.sympify(d).evalf()) for d in str(text).split(',')]
    if len(vals) != 2:
      return tuple(default)
    l1, l2 = vals
    return (l2, l1) if l1 > l2 else (l1, l2)
  except Exception:
    return tuple(default)


def odd_resolution(res):
  'reference random_number_generator.py:330-334'
  res = int(round(float(res)))
  return res+1 if res % 2 == 0 else res


def point_source_density(density, focal_length, scalar=False):
  '''
  Power density string of a point source -> (sympy expression, first variable name), restating
  PointSourceProxy._rvArgs (reference freecad_elements/point_source.py:277-366): the area element
  |sin(theta)| (finite focal length) or |r| (collimated) is multiplied in for the 2-D variable, and
  r, x, y are substituted by their theta/phi (or r/phi) expressions.
  '''
  f = float(focal_length)
  if np.isfinite(f):
    if np.isclose(f, 0):
      stripped = str(density)
      for w in ('exp', 'arcsin', 'arccos', 'arctan2', 'arctan', 'arccot', 'arsinh', 'arcosh', 'artanh',
                'arcoth', 'DiracDelta', 'Piecewise', 'Heaviside', 'True', 'False'):
        stripped = stripped.replace(w, '')
      for c in 'rxy':
        if c in stripped:
          raise ValueError(f'Variable {c} in power density expression {density} is forbidden if focal length is zero.')
    if not scalar:
      density = '('+str(density)+')*abs(sin(theta))'
    fs = f'{abs(f):.8e}'
    expr = (sy.sympify(density)
            .subs('r', sy.sympify(f'(tan(theta)*{fs})'))
            .subs('x', sy.sympify(f'(tan(theta)*cos(phi)*{fs})'))
            .subs('y', sy.sympify(f'(tan(theta)*sin(phi)*{fs})')))
    return expr, 'theta'
  if not scalar:
    density = '('+str(density)+')*abs(r)'
  if 'theta' in str(density):
    raise ValueError(f'Variable theta in power density expression {density} is forbidden if focal length is infinite.')
  expr = (sy.sympify(density)
          .subs('x', sy.sympify('(r*cos(phi))'))
          .subs('y', sy.sympify('(r*sin(phi))')))
  return expr, 'r'


class SamplerTables:
  '''
  phi_cdf   [n_phi]            normalised marginal CDF of phi on edges linspace(phi_domain, n_phi)
  first_cdf [n_rows, n_first]  normalised conditional CDF of theta|r per phi mid-point row
                               (n_rows = n_phi-1, or 1 when the density does not depend on phi)
  '''
  def __init__(self, phi_cdf, first_cdf, first_domain, phi_domain, first_var):
    self.phi_cdf = np.ascontiguousarray(phi_cdf, dtype=np.float64)
    first_cdf = np.asarray(first_cdf, dtype=np.float64)
    self.first_cdf = np.ascontiguousarray(np.atleast_2d(first_cdf) if first_cdf.ndim < 3 else first_cdf)
    self.first_domain = (float(first_domain[0]), float(first_domain[1]))
    self.phi_domain = (float(phi_domain[0]), float(phi_domain[1]))
    self.first_var = first_var

  @property
  def n_tables(self):
    '''
    > 1: a FAMILY of tables over the incidence angle of a hit (scatter_tables with a density that depends on theta_in /
    theta_refl): phi_cdf [n_tables, n_phi], first_cdf [n_tables, n_rows, n_first]; see include/odw.h odw_scatter
    '''
    return self.first_cdf.shape[0] if self.first_cdf.ndim == 3 else 1

  @property
  def n_rows(self):
    return self.first_cdf.shape[-2]

  def table(self, k):
    'member k of a family as a plain SamplerTables'
    if self.n_tables == 1:
      return self
    return SamplerTables(self.phi_cdf[k], self.first_cdf[k], self.first_domain, self.phi_domain, self.first_var)

  # numpy restatement of the draw (random_number_generator.py:413-456,492-500); used by tests
  def draw_from_uniforms(self, u_phi, u_first):
    e_phi = np.linspace(*self.phi_domain, len(self.phi_cdf))
    e_first = np.linspace(*self.first_domain, self.first_cdf.shape[1])
    phi = np.interp(u_phi, self.phi_cdf, e_phi)
    if self.n_rows == 1:
      return np.interp(u_first, self.first_cdf[0], e_first), phi
    c_phi = (e_phi[1:]+e_phi[:-1])/2
    first = np.empty_like(phi)
    for i, (p, u) in enumerate(zip(phi, u_first)):
      row = int(np.argmin(np.abs(c_phi-p)))
      first[i] = np.interp(u, self.first_cdf[row], e_first)
    return first, phi


def build_tables(expr, first_var, first_domain, phi_domain, first_resolution, phi_resolution):
  '''
  expr: sympy expression in (first_var, phi).  Follows _generateNumericScalarLambda: edges
  linspace(l1, l2, odd res), density evaluated at the cell mid-points on meshgrid(C_first, C_phi)
  (shape [n_phi-1, n_first-1]), conditional table = [0, cumsum along first], marginal =
  [0, cumsum of the row sums]; every CDF divided by its last entry.
  '''
  expr = sy.sympify(expr)
  n_first = odd_resolution(first_resolution)
  n_phi = odd_resolution(phi_resolution)
  e_first = np.linspace(first_domain[0], first_domain[1], n_first)
  e_phi = np.linspace(phi_domain[0], phi_domain[1], n_phi)
  c_first = (e_first[1:]+e_first[:-1])/2
  c_phi = (e_phi[1:]+e_phi[:-1])/2
  names = {str(s) for s in expr.free_symbols}
  extra = names-{first_var, 'phi'}
  if extra:
    raise ValueError(f'probability density expression {expr} has free symbols {sorted(extra)} besides {first_var}, phi')
  v1, v2 = sy.Symbol(first_var), sy.Symbol('phi')
  expr = expr.subs({s: (v1 if str(s) == first_var else v2) for s in expr.free_symbols})
  lam = sy.lambdify([v1, v2], expr, modules=['numpy', 'scipy'])
  phi_dependent = 'phi' in names
  if phi_dependent:
    g1, g2 = np.meshgrid(c_first, c_phi)
    probs = np.asarray(lam(g1, g2), dtype=np.float64)
    if probs.shape != g1.shape:
      probs = g1*0+probs
  else:
    probs = np.asarray(lam(c_first, c_phi[0]), dtype=np.float64)
    if probs.shape != c_first.shape:
      probs = c_first*0+probs
    probs = probs[None, :]
  if not np.all(np.isfinite(probs)):
    raise ValueError(f'probability density {expr} is not finite on its domain')
  if (probs < 0).any():
    raise ValueError(f'found negative probability density, expression: {expr}')
  cond = np.cumsum(np.insert(probs, 0, 0.0, axis=-1), axis=-1)
  rowsum = probs.sum(axis=-1)
  if not phi_dependent:
    rowsum = np.full(n_phi-1, rowsum[0])
  marg = np.cumsum(np.insert(rowsum, 0, 0.0))
  cond = cond/cond[:, -1:]
  marg = marg/marg[-1]
  return SamplerTables(marg, cond, first_domain, phi_domain, first_var)


def point_source_tables(rec):
  'source record (scene_export.fcstd.source_records) -> SamplerTables'
  f = float(rec.get('FocalLength', '0'))
  expr, var = point_source_density(rec['PowerDensity'], f)
  if var == 'theta':
    dom = parse_domain(rec.get('ThetaDomain', '0, pi/4'), (0, np.pi/4))
    res = rec.get('ThetaResolutionNumericMode', '1e5')
  else:
    dom = parse_domain(rec.get('RadiusDomain', '0, 10'), (0, 10))
    res = rec.get('RadiusResolutionNumericMode', '1e5')
  phi_dom = parse_domain(rec.get('PhiDomain', '0, 2*pi'), (0, 2*np.pi))
  return build_tables(expr, var, dom, phi_dom, float(res), float(rec.get('PhiResolutionNumericMode', '1e2')))


def draw_pseudo(tables, density, N, rng, overdraw_factor=0.1, overdraw_iterations=50, bins=None):
  '''
  Pseudo-random draws: N samples whose histogram follows the density more closely than N true random draws do
  (reference distributions/random_number_generator.py:562-682 drawPseudo).  Start from (1 + f) N true draws; per round add
  f N fresh draws, histogram the pool, and repeatedly delete one random sample from the bin that exceeds its expected share
  the most until only N are left (or that bin is empty); after `overdraw_iterations` rounds return the last N kept.

  tables   SamplerTables (true draws come from them);  density  sympy expression in (first variable, phi) used for the
  expected histogram (the same expression the tables were built from);  rng  numpy Generator.
  Returns (first, phi) arrays of length N.
  '''
  N = int(round(N))
  if N <= 1:
    raise ValueError('N must be greater than one in pseudo random mode')
  expr = sy.sympify(density)
  v1, v2 = sy.Symbol(tables.first_var), sy.Symbol('phi')
  expected_of = sy.lambdify([v2, v1], expr, modules=['numpy', 'scipy'])       # reversed variable order like the reference
  draw = lambda n: np.stack(tables.draw_from_uniforms(rng.random(n), rng.random(n)))      # rows: first, phi
  if bins is None:
    bins = max(1, int((overdraw_factor*np.sqrt(overdraw_iterations)*N)**(1/(3*2))))
  pool = draw(int(round(N*(1+overdraw_factor))))
  for it in range(int(round(overdraw_iterations))):
    if it > 0:
      pool = np.concatenate([pool[:, ~np.isnan(pool[0])], draw(int(round(N*overdraw_factor)))], axis=-1)
    hist, edges = np.histogramdd(pool.T, bins=bins)
    centres = [(e[1:]+e[:-1])/2 for e in edges]
    expected = expected_of(*np.meshgrid(*reversed(centres)))
    if not hasattr(expected, 'shape') or np.shape(expected) != hist.shape:
      expected = expected*np.ones(hist.shape)
    while True:
      excess = hist/hist.sum()-expected/expected.sum()
      worst = np.argwhere(excess == excess.max())[0]
      inside = np.ones(pool.shape[1], dtype=bool)
      for axis, k in enumerate(worst):
        inside &= (edges[axis][k] < pool[axis]) & (pool[axis] <= edges[axis][k+1])
      members = np.nonzero(inside)[0]
      if len(members) == 0:
        pool = pool[:, ~np.isnan(pool[0])][:, -N:]
        break
      pool[:, members[int(rng.random()*len(members))]] = np.nan
      hist[tuple(worst)] -= 1
      if np.count_nonzero(~np.isnan(pool[0])) <= N:
        break
  result = pool[:, ~np.isnan(pool[0])][:, -N:]
  return result[0], result[1]


def surface_source_tables(rec):
  '''
  Surface source record -> SamplerTables whose single conditional row is the theta CDF.  The reference builds a
  ScalarRandomVariable from the power density as it stands — no sin(theta) area factor (surface_source.py:530 with
  point_source.py:277-321, scalarRandomVar=True; a surface source has no FocalLength, so the finite-focal-length branch
  with f = 1 applies).  phi is uniform in [0, 2 pi) (surface_source.py:544) and needs no table.
  '''
  expr, var = point_source_density(rec['PowerDensity'], 1.0, scalar=True)
  dom = parse_domain(rec.get('ThetaDomain', '0, pi/2'), (0, np.pi/2))
  return build_tables(expr, var, dom, (0.0, 2*np.pi), float(rec.get('ThetaResolutionNumericMode', '1e5')), 3)


SCATTER_PARAM_TABLES = 91        # members of a per-hit family: incidence angles linspace(0, pi/2, 91), one per degree


def specular_angle(theta_in, optical_type, entering, refractive_index):
  '''
  theta_refl of applyStochasticRayCorrections (optical_group.py:292): angle between the IDEAL outgoing direction and the
  face normal flipped along the propagation, as a function of the incidence angle.  Mirror: pi - theta_in.  Lens: the
  refraction angle (vacuum outside; n1 -> n2 = 1 -> n entering, n -> 1 leaving), pi - theta_in beyond the critical angle.
  '''
  if optical_type == 'Mirror':
    return np.pi-theta_in
  mu = 1.0/refractive_index if entering else refractive_index
  s = mu*np.sin(theta_in)
  return float(np.arcsin(s)) if s < 1 else np.pi-theta_in


def scatter_tables(density, theta_domain, phi_domain, resolution=None, optical_type='Mirror', refractive_index=1.0,
                   param_tables=None):
  '''
  Stochastic surface model of an optical group (reference freecad_elements/optical_group.py:212-269): the density string
  in (theta, phi) as it stands (no sin(theta) factor, :219-223) -> SamplerTables, or None when the density is empty or
  reduces to "no change".  Resolution: the reference's default for a 2-variable sampler, 5 + int(1e6**(1/2)) = 1005
  (random_number_generator.py:323-334).

  Handled specially: `DiracDelta(theta)` (optionally times DiracDelta(phi) or a function of phi) pins theta to 0, for
  which both rotations of applyStochasticRayCorrections are the identity (:311-320) -> None.
  Densities that depend on the incident / specular angles (theta_in, phi_in, theta_refl, phi_refl): the reference
  re-compiles them for every hit with theta_in = angle(direction, normal), theta_refl = angle(ideal outgoing direction,
  normal), phi_in = phi_refl = 0 (:288-307).  Here they become a FAMILY of tables over theta_in on linspace(0, pi/2,
  SCATTER_PARAM_TABLES) — theta_refl follows from theta_in (specular_angle; a Lens gets two families, entering and
  leaving) — and a hit uses the member nearest to its incidence angle: the density is evaluated at most half a degree
  off the hit's own angle (documented approximation; the reference evaluates it exactly).  Such densities must not depend
  on phi (one conditional row), or the family would not fit.
  Refused: DiracDelta terms other than DiracDelta(theta).
  '''
  if density is None or not str(density).strip():
    return None
  expr = sy.sympify(str(density))
  names = {str(x) for x in expr.free_symbols}
  per_hit = names & {'theta_in', 'phi_in', 'theta_refl', 'phi_refl'}
  if per_hit:
    if expr.has(sy.DiracDelta):
      raise NotImplementedError(f'stochastic surface density {density!r}: DiracDelta terms with per-hit parameters')
    if 'phi' in names:
      raise NotImplementedError(f'stochastic surface density {density!r} depends on {sorted(per_hit)} AND on phi: the family of '
                                f'per-hit tables would need a conditional table per incidence angle')
    K = int(param_tables or SCATTER_PARAM_TABLES)
    res = 5+int(1e6**0.5) if resolution is None else resolution
    families = [(True,)] if optical_type == 'Mirror' else [(True,), (False,)]
    phis, firsts, one = [], [], None
    for (entering,) in families:
      for th_in in np.linspace(0.0, np.pi/2, K):
        sub = {sy.Symbol('theta_in'): th_in, sy.Symbol('phi_in'): 0.0, sy.Symbol('phi_refl'): 0.0,
               sy.Symbol('theta_refl'): specular_angle(float(th_in), optical_type, entering, float(refractive_index))}
        one = build_tables(expr.subs(sub), 'theta', parse_domain(theta_domain, (-np.pi/2, np.pi/2)),
                           parse_domain(phi_domain, (0, 2*np.pi)), res, 3)
        phis.append(one.phi_cdf); firsts.append(one.first_cdf)
    return SamplerTables(np.stack(phis), np.stack(firsts), one.first_domain, one.phi_domain, 'theta')
  if expr.has(sy.DiracDelta):
    theta = sy.Symbol('theta')
    pinned = [d for d in expr.atoms(sy.DiracDelta) if d.args[0] == theta]
    if pinned:
      return None
    raise NotImplementedError(f'stochastic surface density {density!r}: only DiracDelta(theta) (no change) is supported')
  res = 5+int(1e6**0.5) if resolution is None else resolution
  return build_tables(expr, 'theta', parse_domain(theta_domain, (-np.pi/2, np.pi/2)),
                      parse_domain(phi_domain, (0, 2*np.pi)), res, res)


# ------------------------------------------------------------------------------------------
# deterministic fan grid

def points_with_given_density_1d(X, Y, N):
  'reference distributions/points_by_density.py:25-38'
  X = np.asarray(X, dtype=float)
  Y = np.asarray(Y, dtype=float)
  Xi = np.concatenate([[X[0]-(X[1]-X[0])/2], (X[:-1]+X[1:])/2, [X[-1]+(X[-1]-X[-2])/2]])
  Yi = np.concatenate([[0], np.cumsum(Y)])
  Yi = (Yi-Yi.min())/(Yi.max()-Yi.min())
  Ypick = np.linspace(0, 1, int(round(N)))[1:-1]
  return np.concatenate([[X[0]], np.interp(Ypick, Yi, Xi), [X[-1]]])


def find_grid(expr, var, domain, resolution, N, constants=None):
  '''
  ScalarRandomVariable.findGrid (reference random_number_generator.py:685-725): evaluate the density on
  linspace(domain, odd res) and place N points by quantiles; result clipped to the domain.
  '''
  expr = sy.sympify(expr)
  for k, v in (constants or {}).items():
    if k in [str(s) for s in expr.free_symbols]:
      expr = expr.subs(k, v)
  free = [s for s in expr.free_symbols]
  if len(free) > 1:
    raise ValueError(f'expression "{expr}" seems to have more than one free variable after substituting constants')
  sym = free[0] if free else sy.Symbol(var)
  rng = np.linspace(domain[0], domain[1], odd_resolution(resolution))
  density = sy.lambdify(sym, expr, modules=['numpy', 'scipy'])(rng)
  if not hasattr(density, 'shape') or np.shape(density) != rng.shape:
    density = density*np.ones(rng.shape)
  result = points_with_given_density_1d(rng, density, N)
  return result[np.logical_and(rng.min() <= result, result <= rng.max())]
