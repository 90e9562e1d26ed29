'''Worker of tests/test_gpu_simulation.py: one GPU rank of an NCCL run (sharded trace + device histogram all-reduce).'''
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
  base, scene, n_total = sys.argv[1], sys.argv[2], int(sys.argv[3])
  local = int(os.environ['LOCAL_RANK'])
  torch.cuda.set_device(local)
  dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  rank, world = dist.get_rank(), dist.get_world_size()
  from freecad.optics_design_workbench_b200 import engine
  from freecad.optics_design_workbench_b200.simulation import simulation_loop, sharding
  from freecad.optics_design_workbench_b200.simulation.setup import prepare
  sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', scene+'.npz'))
  eng = engine.Engine(local)
  run = simulation_loop.runSimulation(sim, 'true', engine=eng, basePath=base,
                                      settings=dict(EndAfterRays=n_total-1, RaysPerIteration=n_total//4), maxBatchRays=n_total//2)
  binning = dict(group=len(sim.scene.groups)-1, nu=32, nv=32, origin=(0, 0, 0), uaxis=(1, 0, 0), vaxis=(0, 1, 0),
                 u_range=(-100, 100), v_range=(-100, 100))
  cfg = sim.cfg(store_hits=False, binnings=[binning])
  ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
  first, n = sharding.shard_range(0, n_total, rank, world)
  with ds.trace_mc(dsrc, cfg, simulation_loop.DEFAULT_SEED, first, n) as res:
    ptr, nb = res.histogram_device(0)
    sharding.all_reduce_histogram_device(ptr, nb, local)        # NCCL, in place on the engine's bins
    hist = res.histogram(0)
    counts = res.counts
  total = sharding.all_reduce_counters(dict(segments=counts['segments'], hits=counts['hits'], rays=n))
  with open(f'{base}/rank{rank}.json', 'w') as f:
    json.dump(dict(run=run, rank=rank, world=world, first=first, n=n, hist=hist.tolist(), counters=total), f)
  dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
