'''
Generates tests/golden/sampler_golden.npz by IMPORTING the reference's `distributions` module from
/root/reference (shim from SURVEY.md Appendix E: stub matplotlib, skip the package __init__).  Run in the
build container only (the reference tree does not exist on the GPU box); the .npz is committed.

For every case the reference's numpy RNG is replaced by a recorded uniform stream, so that
  (theta|r, phi) = VectorRandomVariable.draw(N)   (reference distributions/random_number_generator.py:467-560)
is a deterministic function of the recorded uniforms; oracle and CUDA sampler must reproduce it.
Also stores ScalarRandomVariable.findGrid fan grids (random_number_generator.py:685-725).
'''
import sys, types, os
import numpy as np

def load_reference_distributions():
  mpl = types.ModuleType('matplotlib'); plt = types.ModuleType('matplotlib.pyplot'); plt.__all__ = []
  mpl.pyplot = plt
  sys.modules.update({'matplotlib': mpl, 'matplotlib.pyplot': plt})
  root = '/root/reference/freecad/optics_design_workbench'
  pkg = types.ModuleType('odw_ref'); pkg.__path__ = [root]; sys.modules['odw_ref'] = pkg
  sim = types.ModuleType('odw_ref.simulation'); sim.__path__ = []; sys.modules['odw_ref.simulation'] = sim
  proc = types.ModuleType('odw_ref.simulation.processes'); proc.isMasterProcess = lambda: None
  sim.processes = proc; sys.modules['odw_ref.simulation.processes'] = proc
  rs = types.ModuleType('odw_ref.simulation.results_store')     # io._getLogDir(): "no document open" -> no log file
  def _no_doc():
    raise RuntimeError('no FCStd file opened')
  rs.getResultsFolderPath = _no_doc
  sim.results_store = rs; sys.modules['odw_ref.simulation.results_store'] = rs
  from odw_ref import distributions
  return distributions

CASES = [
  # name, density (already including the |sin(theta)| area element like point_source.py:299), first var, domains, resolutions
  dict(name='gauss_minimal', expr='(exp(-theta**2/(1e-2)**2))*abs(sin(theta))', var='theta',
       dom=(0, np.pi/4), phi=(0, 2*np.pi), res=(1e5, 1e2)),
  dict(name='gauss_huge', expr='(exp(-theta**2/(.2)**2))*abs(sin(theta))', var='theta',
       dom=(0, np.pi/4), phi=(0, 2*np.pi), res=(1e4, 1e2)),
  dict(name='astigmatic', expr='(exp(-theta**2/0.05)*(1+0.8*cos(phi)**2))*abs(sin(theta))', var='theta',
       dom=(0, np.pi/3), phi=(0.3, 5.1), res=(2001, 41)),
  dict(name='collimated', expr='(exp(-r**2/4)*(2+sin(phi)))*abs(r)', var='r',
       dom=(0, 6), phi=(0, 2*np.pi), res=(1001, 31)),
  dict(name='flat_zero_tail', expr='(Piecewise((1, theta<0.2), (0, True)))*abs(sin(theta))', var='theta',
       dom=(0, 0.5), phi=(0, 2*np.pi), res=(501, 11)),
]

def main():
  dist = load_reference_distributions()
  rng = np.random.default_rng(20261018)
  out = {}
  N = 4000
  for c in CASES:
    vrv = dist.VectorRandomVariable(c['expr'], variableOrder=(c['var'], 'phi'),
                                    variableDomains={c['var']: c['dom'], 'phi': c['phi']},
                                    numericalResolutions={c['var']: c['res'][0], 'phi': c['res'][1]})
    vrv.compile(disableAnalytical=True)
    assert vrv.mode() == 'numeric'
    u_phi, u_first = rng.random(N), rng.random(N)
    stream = [u_phi, rng.random(N), u_first, rng.random(N)]     # draw order: phi, (discrete-event roll), first, (roll)
    import numpy.random as npr
    orig = npr.random_sample
    dist.random_number_generator.random.random_sample = lambda size=None: stream.pop(0)
    try:
      first, phi = vrv.draw(N=N)
    finally:
      dist.random_number_generator.random.random_sample = orig
    assert not stream
    for k, v in dict(u_phi=u_phi, u_first=u_first, first=first, phi=phi).items():
      out[f"{c['name']}/{k}"] = np.asarray(v, dtype=np.float64)
    out[f"{c['name']}/meta"] = np.array([c['dom'][0], c['dom'][1], c['phi'][0], c['phi'][1], c['res'][0], c['res'][1]])
    out[f"{c['name']}/expr"] = np.array(c['expr'])
    out[f"{c['name']}/var"] = np.array(c['var'])
  # fan grids
  for name, expr, var, dom, res, Ns in [
      ('fan_gauss', 'exp(-theta**2/0.01)', 'theta', (-np.pi/4, np.pi/4), 1e5, (10, 20, 51)),
      ('fan_cos', 'cos(theta)**2', 'theta', (0.1, 1.2), 1e3, (4, 25)),
      ('fan_const', '1', 'r', (0, 10), 101, (5, 16))]:
    for n in Ns:
      srv = dist.ScalarRandomVariable(expr, variableDomain=dom, variable=var, numericalResolution=res)
      srv.compile()
      out[f'{name}/N{n}'] = np.asarray(srv.findGrid(N=n), dtype=np.float64)
    out[f'{name}/meta'] = np.array([dom[0], dom[1], res])
    out[f'{name}/expr'] = np.array(expr)
    out[f'{name}/var'] = np.array(var)
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sampler_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, len(out), 'arrays')

if __name__ == '__main__':
  main()
