'''
GPU: the host-side simulation path (runSimulation -> runSimulationIteration -> C ABI -> hit writer) against the same
path run with the oracle-backed engine, and the multi-GPU NCCL histogram all-reduce (needs >= 2 GPUs, else skipped).
'''
import glob
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from freecad.optics_design_workbench_b200.simulation import simulation_loop
from freecad.optics_design_workbench_b200.simulation.setup import prepare

from conftest import ROOT, SCENES
from oracle_engine import OracleEngine
from test_simulation_loop import load_hits

pytestmark = pytest.mark.gpu


def _sorted(h):
  order = np.lexsort(np.round(np.c_[h['points'], h['directions']], 7).T)
  return {k: (v[order] if isinstance(v, np.ndarray) else v) for k, v in h.items()}


@pytest.mark.parametrize('scene,action,settings', [
  ('minimal', 'true', dict(EndAfterRays=5000)),
  ('lensesAndMirrors', 'true', dict(EndAfterRays=20000, RaysPerIteration=1000)),
  ('lensesAndMirrorsSequential', 'true', dict(EndAfterHits=3000)),
  ('lensesAndMirrors', 'fans', {}),
  ('lensesAndMirrors', 'singletrue', {}),
])
def test_simulation_run_matches_oracle_backed_run(tmp_path, gpu_engine, scene, action, settings):
  keys = ['InitPoint', 'InitDirection', 'InitTheta', 'InitPhi', 'RayIndex', 'FanIndex']
  out = {}
  for label, eng in (('gpu', gpu_engine), ('cpu', OracleEngine())):
    sim = prepare(os.path.join(SCENES, scene+'.npz'))
    sim.settings['store_hit_keys'] = keys
    run = simulation_loop.runSimulation(sim, action, engine=eng, basePath=str(tmp_path/f'{label}.OpticsDesign'),
                                        settings=dict(settings), maxBatchRays=1 << 14)
    out[label] = _sorted(load_hits(run))
  g, c = out['gpu'], out['cpu']
  assert set(g) == set(c)
  assert len(g['points']) == len(c['points']) > 0
  assert g['source'] == c['source'] and g['obj'] == c['obj']
  np.testing.assert_allclose(g['points'], c['points'], rtol=0, atol=1e-9)
  np.testing.assert_allclose(g['directions'], c['directions'], rtol=0, atol=1e-9)
  np.testing.assert_array_equal(g['isEntering'], c['isEntering'])
  np.testing.assert_allclose(g['powers'], c['powers'], rtol=0, atol=1e-12)
  for k in g:
    if k.startswith('init') or k.endswith('Index'):
      np.testing.assert_allclose(g[k], c[k], rtol=0, atol=1e-9, equal_nan=True)


def test_record_rays_on_the_gpu_equals_oracle_backed_run(tmp_path, gpu_engine):
  import pickle
  out = {}
  for label, eng in (('gpu', gpu_engine), ('cpu', OracleEngine())):
    sim = prepare(os.path.join(SCENES, 'lensesAndMirrors.npz'))
    sim.source_records[0]['RecordRays'] = True
    run = simulation_loop.runSimulation(sim, 'singletrue', engine=eng, basePath=str(tmp_path/f'{label}.OpticsDesign'))
    out[label] = pickle.load(open(glob.glob(f'{run}/source-*/*-rays.pkl')[0], 'rb'))
  assert len(out['gpu']) == len(out['cpu']) == 100
  for g, c in zip(out['gpu'], out['cpu']):
    assert g['media'] == c['media']
    np.testing.assert_allclose(g['points'], c['points'], rtol=0, atol=1e-6)      # escape points: 460 mm lever arm
    np.testing.assert_allclose(g['powers'], c['powers'], rtol=0, atol=1e-12)


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


@pytest.mark.timeout(600)
def test_nccl_histogram_allreduce_and_sharded_run(tmp_path, gpu_engine):
  import torch
  n_gpus = torch.cuda.device_count()
  if n_gpus < 2:
    pytest.skip('needs >= 2 GPUs (run with gpurun --gpus 2)')
  world = 2 if n_gpus < 4 else 4
  scene, n_total = 'lensesAndMirrors', 400000
  base = str(tmp_path/'nccl.OpticsDesign')
  os.makedirs(base)
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
         '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
         os.path.join(ROOT, 'tests', '_dist_worker_gpu.py'), base, scene, str(n_total)]
  p = subprocess.run(cmd, capture_output=True, text=True, timeout=560)
  assert p.returncode == 0, p.stdout[-2000:]+p.stderr[-4000:]
  ranks = [json.load(open(f'{base}/rank{r}.json')) for r in range(world)]
  # single GPU reference of the same global ray range
  sim = prepare(os.path.join(SCENES, scene+'.npz'))
  binning = dict(group=len(sim.scene.groups)-1, nu=32, nv=32, origin=(0, 0, 0), uaxis=(1, 0, 0), vaxis=(0, 1, 0),
                 u_range=(-100, 100), v_range=(-100, 100))
  ds, dsrc = gpu_engine.scene(sim.scene), gpu_engine.source(sim.source_args(0))
  with ds.trace_mc(dsrc, sim.cfg(store_hits=False, binnings=[binning]), simulation_loop.DEFAULT_SEED, 0, n_total) as res:
    want, counts = res.histogram(0), res.counts
  assert want.sum() > 0.9*n_total
  for rk in ranks:
    assert np.array_equal(np.array(rk['hist']), want)              # integer counts in fp64: exact, order-independent
    assert rk['counters']['segments'] == counts['segments'] and rk['counters']['hits'] == counts['hits']
    assert rk['counters']['rays'] == n_total
  # hit files of all ranks together == a single-GPU run of the same settings
  run1 = simulation_loop.runSimulation(sim, 'true', engine=gpu_engine, basePath=str(tmp_path/'one.OpticsDesign'),
                                       settings=dict(EndAfterRays=n_total-1, RaysPerIteration=n_total//4), maxBatchRays=n_total//2)
  h1, hN = _sorted(load_hits(run1)), _sorted(load_hits(ranks[0]['run']))
  assert len(h1['points']) == len(hN['points'])
  np.testing.assert_array_equal(h1['points'], hN['points'])
  pids = {os.path.basename(f).split('-')[1] for f in glob.glob(f"{ranks[0]['run']}/source-*/object-*/*-hits.pkl")}
  assert len(pids) == world
