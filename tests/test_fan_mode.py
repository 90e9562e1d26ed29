'''
Fan mode (deterministic ray lists): our host generator against golden vectors produced by the REFERENCE's own
PointSourceProxy._generateRays(mode='fans') (tests/golden/make_fan_golden.py), plus the _makeRay geometry.
'''
import ast
import os

import numpy as np
import pytest

from freecad.optics_design_workbench_b200.freecad_elements import point_source

GOLD = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'fan_golden.npz'))
CASES = sorted({k.split('/')[0] for k in GOLD.files})


@pytest.mark.parametrize('name', CASES)
def test_fan_rays_match_reference_generator(name):
  props = ast.literal_eval(str(GOLD[name+'/props']))
  want = GOLD[name+'/rays']             # fanIndex, rayIndex, theta|r, phi, totalFanCount, totalRaysInFan
  batch = point_source.generate_fan_rays(props, np.eye(4))
  assert len(batch) == len(want)
  md = batch.metadata
  first = md['initTheta'] if np.isfinite(float(props['FocalLength'])) else md['initRadius']
  assert np.array_equal(md['fanIndex'], want[:, 0].astype(int))
  assert np.array_equal(md['rayIndex'], want[:, 1].astype(int))
  assert np.array_equal(md['totalFanCount'], want[:, 4].astype(int))
  assert np.array_equal(md['totalRaysInFan'], want[:, 5].astype(int))
  np.testing.assert_allclose(first, want[:, 2], rtol=0, atol=1e-12)
  np.testing.assert_allclose(md['initPhi'], want[:, 3], rtol=0, atol=1e-12)


def test_make_ray_conventions():
  'SURVEY Q6: phi = 0 points to -y; finite focal length shifts origins so that all rays cross (0,0,f)'
  src = dict(FocalLength='0', Wavelength=632.8)
  b = point_source.make_rays(src, np.eye(4), [0.0, 0.3, 0.3], [0.0, 0.0, np.pi/2])
  np.testing.assert_allclose(b.directions[0], [0, 0, 1], atol=1e-15)
  np.testing.assert_allclose(b.directions[1], [0, -np.sin(0.3), np.cos(0.3)], atol=1e-15)
  np.testing.assert_allclose(b.directions[2], [np.sin(0.3), 0, np.cos(0.3)], atol=1e-15)
  assert np.all(b.origins == 0) and b.wavelength == 632.8
  src = dict(FocalLength='25')
  b = point_source.make_rays(src, np.eye(4), [0.2, 0.4], [0.7, 2.0])
  t = (25-b.origins[:, 2])/b.directions[:, 2]
  np.testing.assert_allclose(b.origins+t[:, None]*b.directions, [[0, 0, 25]]*2, atol=1e-12)
  src = dict(FocalLength='inf')
  b = point_source.make_rays(src, np.eye(4), [2.0], [np.pi/2])
  np.testing.assert_allclose(b.origins[0], [0, -2, 0], atol=1e-15)
  np.testing.assert_allclose(b.directions[0], [0, 0, 1], atol=1e-15)
  assert np.isnan(b.metadata['initTheta'][0]) and b.metadata['initRadius'][0] == 2.0


def test_make_ray_applies_global_placement():
  M = np.eye(4)
  M[:3, :3] = [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]      # z axis -> +x ... rotation about y
  M[:3, 3] = [1, 2, 3]
  b = point_source.make_rays(dict(FocalLength='0'), M, [0.0], [0.0])
  np.testing.assert_allclose(b.origins[0], [1, 2, 3])
  np.testing.assert_allclose(b.directions[0], [1, 0, 0], atol=1e-15)
