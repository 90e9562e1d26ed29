'''
Multi-GPU sharding of the trace path: one process per GPU, the scene replicated, rays partitioned.

The reference parallelises the same way over CPU worker processes (reference
simulation/processes/simulation_loop.py:450-507: N headless FreeCAD workers, independent RNG seeds,
per-worker result files, progress summed by the master, results_store.py:492-512).  Here a worker is a
GPU rank and the RNG stream is ONE Philox stream indexed by the global ray number, so the union of all
ranks' hits does not depend on the number of GPUs.  Hit lists need no collective (the result tree is
multi-file by design); detector histograms and the progress counters are summed with one all-reduce
(NCCL over NVLink on GPUs; gloo in the CPU tests).
'''

import numpy as np


def rank_and_world():
  'rank / world size of the default torch.distributed group, (0, 1) when not initialised'
  try:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
      return dist.get_rank(), dist.get_world_size()
  except ImportError:
    pass
  return 0, 1


def shard_range(first, n, rank, world):
  '''
  Rank `rank` of `world` traces global ray indices [first+lo, first+hi) of the block [first, first+n):
  contiguous, disjoint, covering, sizes differing by at most one.  Returns (first_ray, n_rays).
  '''
  first, n, rank, world = int(first), int(n), int(rank), int(world)
  if not 0 <= rank < world:
    raise ValueError(f'rank {rank} outside world of {world}')
  lo = (n*rank)//world
  hi = (n*(rank+1))//world
  return first+lo, hi-lo


def all_reduce_counters(counters):
  '''
  Sum a dict of integer progress counters (totalTracedRays, totalRecordedHits, …) over all ranks —
  what the reference's master does by reading every worker's progress file (results_store.py:492-505).
  '''
  rank, world = rank_and_world()
  keys = sorted(k for k, v in counters.items() if not isinstance(v, str))
  if world == 1:
    return dict(counters)
  import torch
  import torch.distributed as dist
  dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
  t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=dev)
  dist.all_reduce(t, op=dist.ReduceOp.SUM)
  out = dict(counters)
  out.update({k: int(v) for k, v in zip(keys, t.tolist())})
  return out


class _DevicePointer:
  'exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy'
  def __init__(self, ptr, n):
    self.__cuda_array_interface__ = dict(shape=(int(n),), typestr='<f8', data=(int(ptr), False), version=3, strides=None)


def all_reduce_histogram_device(ptr, n_bins, device_index):
  '''
  In-place NCCL all-reduce(SUM) of a detector histogram living in the engine's device memory
  (odw_result_histogram_device): the bins never visit the host.  No-op for a single rank.
  '''
  rank, world = rank_and_world()
  if world == 1:
    return
  import torch
  import torch.distributed as dist
  t = torch.as_tensor(_DevicePointer(ptr, n_bins), device=torch.device('cuda', device_index))
  dist.all_reduce(t, op=dist.ReduceOp.SUM)
  torch.cuda.current_stream(device_index).synchronize()


def all_reduce_histogram_host(bins):
  'host-array variant (gloo tests, or histograms already copied out); returns the summed array'
  rank, world = rank_and_world()
  bins = np.ascontiguousarray(bins, dtype=np.float64)
  if world == 1:
    return bins
  import torch
  import torch.distributed as dist
  dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
  t = torch.from_numpy(bins.copy()).to(dev)
  dist.all_reduce(t, op=dist.ReduceOp.SUM)
  return t.cpu().numpy()


def broadcast_object(obj, src=0):
  'rank `src` decides (e.g. the run folder name), everybody gets it'
  rank, world = rank_and_world()
  if world == 1:
    return obj
  import torch.distributed as dist
  box = [obj if rank == src else None]
  dist.broadcast_object_list(box, src=src)
  return box[0]


def barrier():
  rank, world = rank_and_world()
  if world > 1:
    import torch.distributed as dist
    dist.barrier()
