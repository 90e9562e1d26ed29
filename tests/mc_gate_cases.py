'''
Cases of the Monte-Carlo statistical gate (tests/test_mc_gate.py, golden by tests/golden/make_mc_gate_golden.py): the point
source of the reference's test/70-point-source-slow (1-test-monte-carlo.ipynb cells 2 and 10: power densities x theta /
radius domains, FocalLength 0 and inf) in front of a spherical detector of radius 100 mm, 1e5 rays each, histogrammed like
the notebook does (cartesian 30 x 30, polar 3 x 50).
'''
import numpy as np

from freecad.optics_design_workbench_b200.scene_export import primitives as prim
from freecad.optics_design_workbench_b200.scene_export.scene import SceneBuilder
from freecad.optics_design_workbench_b200.simulation.setup import PreparedSimulation

N_RAYS = 100000
EXTENT = 10.5        # mm: theta <= 0.1 rad at 100 mm and r <= 10 mm both stay inside
CASES = {
  # name: (PowerDensity, FocalLength, ThetaDomain | RadiusDomain)
  'gauss_0.03/0..0.1': ('exp(-theta**2/0.03**2)', '0', '0, .1'),
  'gauss_0.03/-0.1..0.1': ('exp(-theta**2/0.03**2)', '0', '-.1, .1'),
  'cos30/0..0.1': ('cos(30*theta)**2', '0', '0, .1'),
  'wedge/-0.1..0.1': ('2-abs(theta)', '0', '-.1, .1'),
  'flat/-0.02..-0.01': ('1', '0', '-.02, -.01'),
  'gauss_r3/0..10': ('exp(-r**2/3**2)', 'inf', '0, 10'),
  'cos_r3/-10..10': ('cos(r/3)**2', 'inf', '-10, 10'),
  'wedge_r/-2..-1': ('10-abs(r)', 'inf', '-2, -1'),
}


def scene():
  b = SceneBuilder()
  det = b.add_group('Detector', 'Detector', optical_type='Absorber', record_hits=True)
  b.add_shape(det, prim.sphere(100.0), np.eye(4))
  return b.build()


def source_record(name):
  density, focal, domain = CASES[name]
  rec = dict(name='OpticalPointSource', label='OpticalPointSource', proxy='PointSourceProxy', source_id=0, gpM=np.eye(4), ignored=[],
             PowerDensity=density, Wavelength=500.0, FocalLength=focal, ThetaDomain='0, pi/4', PhiDomain='0, 2*pi', RadiusDomain='0, 10',
             ThetaResolutionNumericMode='1e5', RadiusResolutionNumericMode='1e5', PhiResolutionNumericMode='1e2',
             Fans=2, FanPhi0='0', RaysPerFan=20, FanModePowerSpan=0.9, RaysPerIterationScale=1.0, MaxIntersectionsScale=1.0,
             MaxRayLengthScale=1.0, RecordRays=False, FanModeRayCount=100)
  rec['RadiusDomain' if focal == 'inf' else 'ThetaDomain'] = domain
  return rec


def simulation(name):
  settings = dict(MaxRayLength=1000.0, MaxIntersections=100.0, DistanceTolerance=1e-6, RaysPerIteration=100.0, SequentialMode=False,
                  EndAfterRays=np.inf, EndAfterHits=np.inf, EndAfterIterations=np.inf, store_hit_keys=[], name='OpticalSimulationSettings')
  return PreparedSimulation(scene(), settings, [source_record(name)])


def histograms(points):
  '''
  detector histograms of hit points like Hits.histogram of the reference (jupyter_utils/histogram.py:54,78-85), with
  FIXED bin edges so that two samples can be compared: cartesian 30 x 30 over [-EXTENT, EXTENT]^2, polar 3 azimuth
  sectors x 50 radial bins over [0, EXTENT]
  '''
  x, y = points[:, 0], points[:, 1]
  edges = np.linspace(-EXTENT, EXTENT, 31)
  cart, _, _ = np.histogram2d(x, y, bins=(edges, edges))
  phi, r = np.arctan2(y, x) % (2*np.pi), np.hypot(x, y)
  polar, _, _ = np.histogram2d(phi, r, bins=(np.linspace(0, 2*np.pi, 4), np.linspace(0, EXTENT, 51)))
  return cart, polar


def chi_square_p(a, b, min_count=10):
  '''
  two-sample chi-square test of two histograms with (nearly) equal totals: sum (a-b)^2/(a+b) over the bins with at least
  min_count entries in a+b, the rest pooled into one bin; returns (p value, statistic, degrees of freedom)
  '''
  from scipy import stats
  a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
  big = (a+b) >= min_count
  aa, bb = list(a[big]), list(b[big])
  if (~big).any():
    aa.append(a[~big].sum()); bb.append(b[~big].sum())
  aa, bb = np.array(aa), np.array(bb)
  keep = (aa+bb) > 0
  aa, bb = aa[keep], bb[keep]
  k1, k2 = np.sqrt(bb.sum()/aa.sum()), np.sqrt(aa.sum()/bb.sum())       # unequal totals (Numerical Recipes chstwo)
  stat = float(((k1*aa-k2*bb)**2/(aa+bb)).sum())
  dof = len(aa)-1
  return float(stats.chi2.sf(stat, dof)), stat, dof
