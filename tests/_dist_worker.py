'''Worker of tests/test_sharding.py: one rank of a world_size-2 gloo run of the sharded simulation loop (CPU).'''
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch.distributed as dist


def main():
  base, scene = sys.argv[1], sys.argv[2]
  dist.init_process_group('gloo')
  rank, world = dist.get_rank(), dist.get_world_size()
  from freecad.optics_design_workbench_b200.simulation import simulation_loop, sharding
  from freecad.optics_design_workbench_b200.simulation.setup import prepare
  from oracle_engine import OracleEngine
  from oracle import Oracle
  sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', scene+'.npz'))
  run = simulation_loop.runSimulation(sim, 'true', engine=OracleEngine(), basePath=base,
                                      settings=dict(EndAfterRays=3000, RaysPerIteration=500), maxBatchRays=1500)
  # detector histogram: every rank bins its shard of rays [0, 4000), then one all-reduce
  binning = dict(group=len(sim.scene.groups)-1, nu=8, nv=8, origin=(0, 0, 0), uaxis=(1, 0, 0), vaxis=(0, 1, 0),
                 u_range=(-100, 100), v_range=(-100, 100))
  cfg = sim.cfg(store_hits=False, binnings=[binning])
  first, n = sharding.shard_range(0, 4000, rank, world)
  r = Oracle().trace_mc(sim.scene, sim.source_args(0), cfg, simulation_loop.DEFAULT_SEED, first, n)
  total = sharding.all_reduce_histogram_host(r['histograms'][0])
  counters = sharding.all_reduce_counters(dict(segments=r['counts']['segments'], rays=n, label='x'))
  with open(f'{base}/rank{rank}.json', 'w') as f:
    json.dump(dict(run=run, rank=rank, world=world, first=first, n=n, hist=total.tolist(), counters=counters), f)
  dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
