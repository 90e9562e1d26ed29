'''
Pins the sampler restatements (numpy table builder + C oracle) and the fan grid against golden vectors
generated from the reference's own `distributions` module (tests/golden/make_sampler_golden.py).
'''
import os

import numpy as np
import pytest

from freecad.optics_design_workbench_b200 import _abi
from freecad.optics_design_workbench_b200.distributions import build_tables, find_grid, point_source_density

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'sampler_golden.npz'))
DRAW_CASES = sorted({k.split('/')[0] for k in GOLD.files if not k.startswith('fan')})
FAN_CASES = sorted({k.split('/')[0] for k in GOLD.files if k.startswith('fan')})


def _tables(name):
  meta = GOLD[name+'/meta']
  return build_tables(str(GOLD[name+'/expr']), str(GOLD[name+'/var']), (meta[0], meta[1]), (meta[2], meta[3]),
                      meta[4], meta[5])


@pytest.mark.parametrize('name', DRAW_CASES)
def test_numpy_tables_reproduce_reference_draws(name):
  t = _tables(name)
  first, phi = t.draw_from_uniforms(GOLD[name+'/u_phi'], GOLD[name+'/u_first'])
  assert np.abs(first-GOLD[name+'/first']).max() < 1e-13
  assert np.abs(phi-GOLD[name+'/phi']).max() < 1e-13


@pytest.mark.parametrize('name', DRAW_CASES)
def test_oracle_sampler_reproduces_reference_draws(name, oracle):
  t = _tables(name)
  sa = _abi.SourceArgs(t, kind=0, source_id=0, gpM=np.eye(4))
  first, phi = oracle.sample_uniforms(sa, GOLD[name+'/u_phi'], GOLD[name+'/u_first'])
  assert np.abs(first-GOLD[name+'/first']).max() < 1e-13
  assert np.abs(phi-GOLD[name+'/phi']).max() < 1e-13


def test_phi_independent_density_collapses_to_one_row():
  assert _tables('gauss_minimal').n_rows == 1
  assert _tables('astigmatic').n_rows == 40


@pytest.mark.parametrize('name', FAN_CASES)
def test_fan_grid_matches_reference(name):
  meta = GOLD[name+'/meta']
  for k in GOLD.files:
    if k.startswith(name+'/N'):
      n = int(k.split('N')[-1])
      mine = find_grid(str(GOLD[name+'/expr']), str(GOLD[name+'/var']), (meta[0], meta[1]), meta[2], n)
      assert len(mine) == len(GOLD[k]) == n
      assert np.abs(mine-GOLD[k]).max() < 1e-13


def test_fan_grid_properties_from_reference_notebook():
  'assertions of test/10-pure-python-notebooks/meshes_by_density.ipynb: len == N, centre exactly 0, mirror symmetry'
  for n in (11, 21, 51):
    g = find_grid('exp(-theta**2/0.01)', 'theta', (-np.pi/4, np.pi/4), 1e5, n)
    assert len(g) == n
    assert abs(g[n//2]) < 1e-9
    assert np.abs(g+g[::-1]).max() < 1e-9


def test_area_element_is_applied_like_point_source():
  expr, var = point_source_density('exp(-theta**2/(1e-2)**2)', 0.0)
  assert var == 'theta' and 'sin' in str(expr)
  expr, var = point_source_density('exp(-r**2)', float('inf'))
  assert var == 'r' and 'Abs(r)' in str(expr)
  with pytest.raises(ValueError):
    point_source_density('exp(-r**2)', 0.0)       # r is forbidden for focal length 0 (point_source.py:286-294)
  with pytest.raises(ValueError):
    point_source_density('exp(-theta**2)', float('inf'))


def test_one_dimensional_histogram_rms_like_reference_notebook(oracle):
  'distributions_quantitative.ipynb cell 15: 1-D histogram RMS < 3e-2 — here for theta of the benchmark source'
  t = _tables('gauss_huge')
  sa = _abi.SourceArgs(t, kind=0, source_id=0, gpM=np.eye(4))
  rng = np.random.default_rng(1)
  n = 400000
  first, phi = oracle.sample_uniforms(sa, rng.random(n), rng.random(n))
  hist, edges = np.histogram(first, bins=40, range=(0, 0.6), density=True)
  c = (edges[1:]+edges[:-1])/2
  expected = np.exp(-c**2/0.2**2)*np.sin(c)
  expected /= np.trapezoid(expected, c)
  assert np.sqrt(np.mean((hist/hist.max()-expected/expected.max())**2)) < 3e-2
  hphi, _ = np.histogram(phi, bins=20, range=(0, 2*np.pi))
  assert np.abs(hphi/hphi.mean()-1).max() < 0.05
