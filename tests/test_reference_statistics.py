'''
The quantitative assertions of the reference's own integration tests, run against our path (oracle on the CPU, CUDA
kernel on the GPU) on the same project files (imported headless into tests/golden/scenes/):

  test/50-old-tests/run-simulations.py:125-174   gaussian.FCStd: >= 0.8e5 hits of 1e5 rays; Gaussian fit of the central
                                                 cross-sections of a 30x30 histogram: sigma within 30 % of
                                                 100 mm * sqrt(1e-4) = 1.0 mm, centre within 0.5 mm
  test/70-point-source-slow (1-test-monte-carlo)  hits on a detector reproduce the source's angular power density
                                                 (here: Kolmogorov-Smirnov against the analytic CDF instead of the
                                                 notebook's binned RMS thresholds)
'''
import numpy as np
import pytest
import scipy.optimize
from scipy import stats

from conftest import SEED


def gaussian_example_checks(points, n_rays):
  assert len(points) > 0.8*n_rays
  Hs, Xs, Ys = np.histogram2d(points[:, 0], points[:, 1], bins=30)
  # Hs[i, j]: x bin i, y bin j (numpy.histogram2d and matplotlib.hist2d agree)
  gaussian = lambda X, A, s, x0: A*np.exp(-(X-x0)**2/s**2)
  distance, theta_sigma = 100, np.sqrt(1e-4)
  for X, Y in (((Xs[1:]+Xs[:-1])/2, Hs[:, np.argmin(np.abs((Ys[1:]+Ys[:-1])/2))]),
               ((Ys[1:]+Ys[:-1])/2, Hs[np.argmin(np.abs((Xs[1:]+Xs[:-1])/2)), :])):
    popt, _ = scipy.optimize.curve_fit(gaussian, X, Y, p0=(max(Y), 10, 0))
    found, theory = abs(popt[1]), distance*theta_sigma
    assert abs(found-theory)/found < 0.3                    # reference assertion
    assert abs(popt[-1]) < 0.5                              # reference assertion


def test_gaussian_example_on_the_oracle(oracle, sims):
  sim = sims('gaussian')
  n = 100000
  r = oracle.trace_mc(sim.scene, sim.source_args(0), sim.cfg(), SEED, 0, n, hit_capacity=2*n, threads=0)
  gaussian_example_checks(r['hits']['points'], n)
  np.testing.assert_allclose(r['hits']['points'][:, 2], 100.0, atol=1e-9)


@pytest.mark.gpu
def test_gaussian_example_on_the_gpu(gpu_engine, sims):
  sim = sims('gaussian')
  n = 100000
  with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sim.source_args(0)), sim.cfg(hit_capacity=2*n), SEED, 0, n) as res:
    h = res.hits(sort=True)
  gaussian_example_checks(h['points'], n)


def angular_density_check(points, n_rays):
  'minimal.FCStd: exp(-theta^2/1e-4) * sin(theta) on a plane at z = 15: theta = atan(rho / 15) must follow the density'
  assert len(points) == n_rays
  theta = np.arctan2(np.hypot(points[:, 0], points[:, 1]), points[:, 2])
  grid = np.linspace(0, np.pi/4, 200001)
  pdf = np.exp(-grid**2/1e-4)*np.abs(np.sin(grid))
  cdf = np.concatenate([[0], np.cumsum((pdf[1:]+pdf[:-1])/2)]); cdf /= cdf[-1]
  assert stats.kstest(theta, lambda x: np.interp(x, grid, cdf)).pvalue > 0.01
  phi = np.arctan2(points[:, 0], -points[:, 1]) % (2*np.pi)        # phi = 0 points to -y
  assert stats.kstest(phi, stats.uniform(0, 2*np.pi).cdf).pvalue > 0.01


def test_detector_hits_follow_the_source_density_on_the_oracle(oracle, sims):
  sim = sims('minimal')
  n = 100000
  r = oracle.trace_mc(sim.scene, sim.source_args(0), sim.cfg(), SEED, 0, n, hit_capacity=2*n, threads=0)
  angular_density_check(r['hits']['points'], n)


@pytest.mark.gpu
def test_detector_hits_follow_the_source_density_on_the_gpu(gpu_engine, sims):
  sim = sims('minimal')
  n = 100000
  with gpu_engine.scene(sim.scene).trace_mc(gpu_engine.source(sim.source_args(0)), sim.cfg(hit_capacity=2*n), SEED, 0, n) as res:
    h = res.hits(sort=True)
  angular_density_check(h['points'], n)
