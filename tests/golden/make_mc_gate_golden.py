'''
Generates tests/golden/mc_gate_golden.npz: detector histograms of the REFERENCE's own Monte-Carlo sampler.
For every case of tests/mc_gate_cases.py the reference's unmodified PointSourceProxy._generateRays(mode='true')
(freecad_elements/point_source.py:659-679: _getVrv -> VectorRandomVariable.compile / draw with numpy's global RNG,
distributions/random_number_generator.py:72-120,467-560; then _makeRay, point_source.py:411-460) produces 1e5 rays under
the FreeCAD stand-ins of tests/freecad_stub.py; the rays are traced to the spherical detector by the oracle (explicit ray
list) and histogrammed like test/70-point-source-slow/1-test-monte-carlo.ipynb does (cartesian 30 x 30, polar 3 x 50).
The north star's gate — chi-square p > 0.01 at matched ray count — compares the device's Philox-driven histograms with these.
Run here only (/root/reference does not exist on the GPU box):  python tests/golden/make_mc_gate_golden.py
'''
import os, sys, types
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]

import freecad_stub
freecad_stub.install()
from freecad_stub import Matrix
from make_fan_golden import load_reference_point_source
import mc_gate_cases as cases
from oracle import Oracle


class Obj:
  'a document object: plain attributes, hashable by identity (raytracing_cache keys on the object)'
  def __init__(self, **kw):
    self.__dict__.update(kw)


def main():
  ps = load_reference_point_source()
  ps.keepGuiResponsiveAndRaiseIfSimulationDone = lambda *a, **k: None
  ps.find = types.SimpleNamespace(activeSimulationSettings=lambda: types.SimpleNamespace(RaysPerIteration=cases.N_RAYS))
  oracle = Oracle()
  out = {}
  np.random.seed(20261018)                                  # the reference draws from numpy's global RNG
  for name in cases.CASES:
    sim = cases.simulation(name)
    rec = sim.source_records[0]
    obj = Obj(**{k: v for k, v in rec.items() if k[:1].isupper()}, Name=rec['name'], Label=rec['label'], RandomNumberGeneratorMode='?')
    proxy = ps.PointSourceProxy.__new__(ps.PointSourceProxy)
    proxy._ensurePropertiesExist = lambda obj: None
    identity = Matrix()
    proxy._getCoordinateTransformMatricesWithoutLinks = lambda obj: (identity, identity, identity, identity)
    rays = list(proxy._generateRays(obj, mode='true'))
    assert len(rays) == cases.N_RAYS and obj.RandomNumberGeneratorMode in ('numeric', 'analytic'), obj.RandomNumberGeneratorMode
    o = np.array([tuple(r.initPoint) for r in rays]); d = np.array([tuple(r.initDirection) for r in rays])
    r = oracle.trace_rays(sim.scene, sim.cfg(), o, d, hit_capacity=2*len(o))
    assert r['rc'] == 0 and len(r['hits']['powers']) == cases.N_RAYS
    cart, polar = cases.histograms(r['hits']['points'])
    out[name+'/cartesian'], out[name+'/polar'] = cart, polar
    out[name+'/mode'] = np.array(obj.RandomNumberGeneratorMode)
    print(name, obj.RandomNumberGeneratorMode, int(cart.sum()), int(polar.sum()), flush=True)
  path = os.path.join(HERE, 'mc_gate_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
