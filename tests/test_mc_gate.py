'''
The north star's Monte-Carlo gate: detector histograms must be statistically consistent with the reference —
chi-square p > 0.01 at matched ray count.  Reference side: tests/golden/mc_gate_golden.npz = histograms of 1e5 rays
drawn by the reference's OWN sampler (PointSourceProxy._generateRays('true') -> VectorRandomVariable.draw with numpy's
RNG, random_number_generator.py:467-560) and traced to the detector (generator tests/golden/make_mc_gate_golden.py).
Our side: the same source traced with the engine's Philox-driven sampler — the oracle on the CPU, the CUDA path through
the C ABI on the GPU.  Scene, densities, domains and binnings are those of the reference's quantitative test
test/70-point-source-slow/1-test-monte-carlo.ipynb (cells 2-7 spherical, cells 10-14 collimated): cartesian 30 x 30 and
polar 3 x 50.  The notebook's own criterion (binned RMS against the analytic density, median < 0.3) is far looser.
'''
import os

import numpy as np
import pytest

import mc_gate_cases as cases
from conftest import SEED

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'mc_gate_golden.npz')
P_MIN = 0.01


@pytest.fixture(scope='module')
def golden():
  z = np.load(GOLDEN)
  return {name: (z[name+'/cartesian'], z[name+'/polar']) for name in cases.CASES}


def gate(points, want, name):
  assert len(points) == cases.N_RAYS                      # every ray reaches the detector sphere
  cart, polar = cases.histograms(points)
  for label, ours, ref in (('cartesian 30x30', cart, want[0]), ('polar 3x50', polar, want[1])):
    assert ours.sum() == ref.sum() == cases.N_RAYS        # matched ray count, nothing outside the binning window
    p, stat, dof = cases.chi_square_p(ours, ref)
    assert p > P_MIN, f'{name}, {label}: chi-square {stat:.1f} at {dof} degrees of freedom, p = {p:.2e}'


def test_the_gate_can_fail():
  'power of the test: a 3 % rescaling of the beam is rejected, two halves of one sample are not'
  rng = np.random.default_rng(3)
  a = rng.normal(0, 3.0, (cases.N_RAYS, 2))
  b = rng.normal(0, 3.0*1.03, (cases.N_RAYS, 2))
  c = rng.normal(0, 3.0, (cases.N_RAYS, 2))
  ha, hb, hc = (cases.histograms(x)[0] for x in (a, b, c))
  assert cases.chi_square_p(ha, hb)[0] < P_MIN
  assert cases.chi_square_p(ha, hc)[0] > P_MIN


@pytest.mark.parametrize('name', list(cases.CASES))
def test_oracle_histograms_are_consistent_with_the_reference_sampler(name, golden, oracle):
  sim = cases.simulation(name)
  r = oracle.trace_mc(sim.scene, sim.source_args(0), sim.cfg(), SEED, 0, cases.N_RAYS, hit_capacity=2*cases.N_RAYS, threads=0)
  gate(r['hits']['points'], golden[name], name)


@pytest.mark.gpu
@pytest.mark.parametrize('name', list(cases.CASES))
def test_gpu_histograms_are_consistent_with_the_reference_sampler(name, golden, gpu_engine):
  sim = cases.simulation(name)
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  with ds.trace_mc(dsrc, sim.cfg(hit_capacity=2*cases.N_RAYS), SEED, 0, cases.N_RAYS) as res:
    h = res.hits(sort=True)
  ds.close(); dsrc.close()
  gate(h['points'], golden[name], name)


@pytest.mark.gpu
def test_gpu_device_binning_is_consistent_with_the_reference_sampler(golden, gpu_engine):
  'the same gate on a histogram binned ON THE DEVICE (odw_binning) instead of from the hit list'
  name = 'gauss_0.03/0..0.1'
  sim = cases.simulation(name)
  e = cases.EXTENT
  spec = dict(group=0, nu=30, nv=30, origin=(0, 0, 0), uaxis=(1, 0, 0), vaxis=(0, 1, 0), u_range=(-e, e), v_range=(-e, e))
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  with ds.trace_mc(dsrc, sim.cfg(store_hits=False, binnings=[spec]), SEED, 0, cases.N_RAYS) as res:
    bins = res.histogram(0)
  ds.close(); dsrc.close()
  assert bins.sum() == cases.N_RAYS
  p, stat, dof = cases.chi_square_p(bins, golden[name][0])
  assert p > P_MIN, (stat, dof, p)
