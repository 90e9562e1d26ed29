'''
Generates tests/golden/fan_golden.npz by running the reference's OWN fan-mode ray generator
(PointSourceProxy._generateRays(mode='fans'), reference freecad_elements/point_source.py:474-656) in the
build container.  FreeCAD is absent, but point_source.py guards its FreeCAD imports, so the module imports
with the package __init__ files bypassed (synthetic parent packages, as in make_sampler_golden.py).  The
only FreeCAD-dependent call on the path, _makeRay (Rotation/Vector arithmetic), is replaced by a recorder:
the golden data are the (fanIndex, rayIndex, theta|r, phi, totalFanCount, totalRaysInFan) tuples in the
order the reference yields them.  Run here only (/root/reference does not exist on the GPU box).
'''
import os, sys, types
import numpy as np


def load_reference_point_source():
  mpl = types.ModuleType('matplotlib'); plt = types.ModuleType('matplotlib.pyplot'); plt.__all__ = []
  mpl.pyplot = plt
  sys.modules.update({'matplotlib': mpl, 'matplotlib.pyplot': plt})
  root = '/root/reference/freecad/optics_design_workbench'
  pkg = types.ModuleType('odw_ref'); pkg.__path__ = [root]; sys.modules['odw_ref'] = pkg
  sim = types.ModuleType('odw_ref.simulation'); sim.__path__ = [root+'/simulation']; sys.modules['odw_ref.simulation'] = sim
  proc = types.ModuleType('odw_ref.simulation.processes'); proc.isMasterProcess = lambda: None
  sim.processes = proc; sys.modules['odw_ref.simulation.processes'] = proc
  rs = types.ModuleType('odw_ref.simulation.results_store')
  def _no_doc():
    raise RuntimeError('no FCStd file opened')
  rs.getResultsFolderPath = _no_doc
  sim.results_store = rs; sys.modules['odw_ref.simulation.results_store'] = rs
  fe = types.ModuleType('odw_ref.freecad_elements'); fe.__path__ = [root+'/freecad_elements']
  sys.modules['odw_ref.freecad_elements'] = fe
  from odw_ref.freecad_elements import point_source
  return point_source


CASES = {
  # shipped defaults of the benchmark sources (stitched fans: theta domain starts at 0)
  'benchmark_default': dict(PowerDensity='exp(-theta**2/(1e-2)**2)', FocalLength='0', ThetaDomain='0, pi/4',
                            PhiDomain='0, 2*pi', Fans=2, RaysPerFan=20, FanPhi0='0', FanModePowerSpan=0.9),
  'test70_three_fans': dict(PowerDensity='exp(-theta^2/0.01)', FocalLength='0', ThetaDomain='0, pi/4',
                            PhiDomain='0, 2*pi', Fans=3, RaysPerFan=50, FanPhi0='0.3', FanModePowerSpan=0.9),
  'gapped': dict(PowerDensity='cos(theta)**2', FocalLength='0', ThetaDomain='0.1, 0.8',
                 PhiDomain='0, 2*pi', Fans=2, RaysPerFan=9, FanPhi0='0', FanModePowerSpan=1.0),
  'theta_sign_change': dict(PowerDensity='exp(-theta**2/0.05)', FocalLength='0', ThetaDomain='-0.3, 0.5',
                            PhiDomain='0, 2*pi', Fans=4, RaysPerFan=11, FanPhi0='pi/8', FanModePowerSpan=0.8),
  'half_phi_domain': dict(PowerDensity='exp(-theta**2/0.02)', FocalLength='0', ThetaDomain='0, 0.6',
                          PhiDomain='0, pi/2', Fans=3, RaysPerFan=12, FanPhi0='0.1', FanModePowerSpan=0.95),
  'phi_dependent': dict(PowerDensity='exp(-theta**2/0.02)*(1+0.5*cos(phi))', FocalLength='0', ThetaDomain='0, 0.6',
                        PhiDomain='0, 2*pi', Fans=3, RaysPerFan=12, FanPhi0='0.1', FanModePowerSpan=0.95),
  'collimated': dict(PowerDensity='exp(-r**2/4)', FocalLength='inf', RadiusDomain='0, 5',
                     PhiDomain='0, 2*pi', Fans=2, RaysPerFan=15, FanPhi0='0', FanModePowerSpan=0.9),
  'focused_astigmatic': dict(PowerDensity='exp(-x**2/2-y**2/8)', FocalLength='25', ThetaDomain='0, pi/6',
                             PhiDomain='0, 2*pi', Fans=2, RaysPerFan=16, FanPhi0='0', FanModePowerSpan=1.0),   # span<1 with x,y raises in the reference
}
DEFAULTS = dict(ThetaResolutionNumericMode='1e5', RadiusResolutionNumericMode='1e5', PhiResolutionNumericMode='1e2',
                RadiusDomain='0, 10', ThetaDomain='0, pi/4', Wavelength=500.0)


def main():
  ps = load_reference_point_source()
  out = {}
  for name, props in CASES.items():
    obj = types.SimpleNamespace(**{**DEFAULTS, **props}, Name='OpticalPointSource', Label='OpticalPointSource')
    proxy = ps.PointSourceProxy.__new__(ps.PointSourceProxy)
    proxy._ensurePropertiesExist = lambda obj: None
    rec = []
    print('case', name, flush=True)
    proxy._makeRay = lambda obj, thetaOrRadius, phi, power=1, metadata={}: rec.append(
      (metadata['fanIndex'], metadata['rayIndex'], thetaOrRadius, phi, metadata['totalFanCount'], metadata['totalRaysInFan']))
    ps.keepGuiResponsiveAndRaiseIfSimulationDone = lambda *a, **k: None
    list(proxy._generateRays(obj, mode='fans'))
    out[name+'/rays'] = np.array(rec, dtype=np.float64)
    out[name+'/props'] = np.array(repr({**DEFAULTS, **props}))
    print(name, len(rec), 'rays')
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'fan_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path)


if __name__ == '__main__':
  main()
