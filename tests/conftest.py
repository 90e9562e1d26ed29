import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

SCENES = os.path.join(ROOT, 'tests', 'golden', 'scenes')
REFERENCE = '/root/reference'
SEED = 0x0DDB1A5E


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
  config.addinivalue_line('markers', 'reference: needs the read-only reference tree at /root/reference')


def pytest_collection_modifyitems(config, items):
  have_ref = os.path.isdir(REFERENCE)
  skip_ref = pytest.mark.skip(reason='/root/reference not present on this machine')
  for item in items:
    if 'reference' in item.keywords and not have_ref:
      item.add_marker(skip_ref)


@pytest.fixture(scope='session')
def oracle():
  from oracle import Oracle
  return Oracle()


@pytest.fixture(scope='session')
def sims():
  from freecad.optics_design_workbench_b200.simulation.setup import prepare
  cache = {}
  def get(name):
    if name not in cache:
      cache[name] = prepare(os.path.join(SCENES, name+'.npz'))
    return cache[name]
  return get


@pytest.fixture(scope='session')
def gpu_engine():
  from freecad.optics_design_workbench_b200 import engine
  eng = engine.Engine(0)       # raises loudly when the library or the device is missing
  yield eng
  eng.close()
