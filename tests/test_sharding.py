'''
Multi-GPU host logic on the CPU: ray-range sharding, and a REAL world_size-2 torch.distributed run (gloo) of the
sharded simulation loop with the oracle-backed engine.  The union of the two ranks' hit files must equal the
single-process run hit for hit (the Philox counter is the global ray index), and the all-reduced histogram must
equal the single-process histogram.
'''
import glob
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from freecad.optics_design_workbench_b200.simulation import sharding, simulation_loop
from freecad.optics_design_workbench_b200.simulation.setup import prepare

from conftest import ROOT, SCENES
from oracle_engine import OracleEngine
from test_simulation_loop import load_hits


def test_shard_range_partitions_exactly():
  for n in (0, 1, 7, 100, 12345):
    for world in (1, 2, 3, 8):
      spans = [sharding.shard_range(1000, n, r, world) for r in range(world)]
      assert spans[0][0] == 1000 and sum(c for _, c in spans) == n
      for (f0, c0), (f1, _) in zip(spans, spans[1:]):
        assert f0+c0 == f1
      assert max(c for _, c in spans)-min(c for _, c in spans) <= 1
  with pytest.raises(ValueError):
    sharding.shard_range(0, 10, 2, 2)


def test_single_process_helpers_are_identity():
  assert sharding.rank_and_world() == (0, 1)
  assert sharding.all_reduce_counters(dict(a=3, s='x')) == dict(a=3, s='x')
  h = np.arange(6.0)
  assert np.array_equal(sharding.all_reduce_histogram_host(h), h)
  assert sharding.broadcast_object('run') == 'run'


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_gloo_run_equals_single_process(tmp_path):
  scene = 'lensesAndMirrors'
  base2 = str(tmp_path/'two.OpticsDesign')
  os.makedirs(base2)
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
         '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
         os.path.join(ROOT, 'tests', '_dist_worker.py'), base2, scene]
  env = dict(os.environ, OMP_NUM_THREADS='2')
  p = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=env)
  assert p.returncode == 0, p.stdout[-2000:]+p.stderr[-4000:]
  ranks = [json.load(open(f'{base2}/rank{r}.json')) for r in range(2)]
  assert ranks[0]['run'] == ranks[1]['run']                      # one run folder, decided by rank 0
  assert ranks[0]['first'] == 0 and ranks[1]['first'] == 2000 and ranks[0]['n'] == ranks[1]['n'] == 2000
  run2 = ranks[0]['run']
  assert len([f for f in os.listdir(run2) if f.startswith('uid-')]) == 1
  pids = {os.path.basename(f).split('-')[1] for f in glob.glob(f'{run2}/source-*/object-*/*-hits.pkl')}
  assert len(pids) == 2                                          # each rank wrote its own files
  # single process, same settings
  sim = prepare(os.path.join(SCENES, scene+'.npz'))
  run1 = simulation_loop.runSimulation(sim, 'true', engine=OracleEngine(), basePath=str(tmp_path/'one.OpticsDesign'),
                                       settings=dict(EndAfterRays=3000, RaysPerIteration=500), maxBatchRays=1500)
  h1, h2 = load_hits(run1), load_hits(run2)
  assert len(h1['points']) == len(h2['points']) > 3000
  key = lambda h: np.lexsort(np.c_[h['points'], h['directions']].T)
  np.testing.assert_array_equal(h1['points'][key(h1)], h2['points'][key(h2)])
  np.testing.assert_array_equal(h1['directions'][key(h1)], h2['directions'][key(h2)])
  # histogram + counters: all-reduced result on both ranks == single-process result
  from oracle import Oracle
  binning = dict(group=len(sim.scene.groups)-1, nu=8, nv=8, origin=(0, 0, 0), uaxis=(1, 0, 0), vaxis=(0, 1, 0),
                 u_range=(-100, 100), v_range=(-100, 100))
  r = Oracle().trace_mc(sim.scene, sim.source_args(0), sim.cfg(store_hits=False, binnings=[binning]),
                        simulation_loop.DEFAULT_SEED, 0, 4000)
  for rk in ranks:
    assert np.array_equal(np.array(rk['hist']), r['histograms'][0])
    assert rk['counters']['segments'] == r['counts']['segments'] and rk['counters']['rays'] == 4000
    assert rk['counters']['label'] == 'x'
  assert r['histograms'][0].sum() > 3000
