'''
BASELINE.json configs[4]: Lambertian surface source, continuous Monte-Carlo, detector histograms binned on the device and
summed over the GPUs with one NCCL all-reduce per step (no hit lists leave the device).

  python tools/bench_histogram.py [--rays 1e8] [--steps 5] [--scene surfaceSourceTest21]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_histogram.py ...

Scene: the reference's test/21-simulation-modes project (surface source Box001.Face5 with cos(theta)**2 -> sphere lens
-> absorber box; fixture tests/golden/scenes/surfaceSourceTest21.npz).  The docs pages BASELINE names for this config
are empty stubs in the reference.  Prints one JSON line on rank 0.
'''
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)

def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--rays', type=float, default=1e8)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=2)
  ap.add_argument('--scene', default='surfaceSourceTest21')
  ap.add_argument('--bins', type=int, default=1000)
  args = ap.parse_args()
  import numpy as np, torch, torch.distributed as dist
  from freecad.optics_design_workbench_b200 import engine
  from freecad.optics_design_workbench_b200.simulation import sharding, simulation_loop
  from freecad.optics_design_workbench_b200.simulation.setup import prepare
  rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', args.scene+'.npz'))
  eng = engine.Engine(local)
  ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
  detector = len(sim.scene.groups)-1                     # the absorber group
  faces = sim.scene.faces[sim.scene.faces['group'] == detector]
  lo, hi = faces['aabb_min'].min(axis=0), faces['aabb_max'].max(axis=0)
  binning = dict(group=detector, nu=args.bins, nv=args.bins, origin=(0, 0, 0), uaxis=(1, 0, 0), vaxis=(0, 1, 0),
                 u_range=(float(lo[0]), float(hi[0])), v_range=(float(lo[1]), float(hi[1])), weighted=1)
  cfg = sim.cfg(store_hits=False, binnings=[binning])
  n = int(args.rays)
  total = None
  def step(k):
    nonlocal total
    first, cnt = sharding.shard_range(k*n*world, n*world, rank, world)
    t0 = time.perf_counter()
    with ds.trace_mc(dsrc, cfg, simulation_loop.DEFAULT_SEED, first, cnt) as res:
      ms = res.kernel_ms
      t1 = time.perf_counter()
      ptr, nb = res.histogram_device(0)
      sharding.all_reduce_histogram_device(ptr, nb, local)           # NCCL, in place, bins never visit the host
      t2 = time.perf_counter()
      c = res.counts
      if k == args.warmup+args.steps-1:
        total = res.histogram(0)
    return c, ms, (t2-t1)*1e3
  for k in range(args.warmup):
    step(k)
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  segs = hits = 0; kms = ams = 0.0
  for k in range(args.warmup, args.warmup+args.steps):
    c, ms, a = step(k)
    segs += c['segments']; hits += c['hits']; kms += ms; ams += a
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  dt = time.perf_counter()-t0
  tot = sharding.all_reduce_counters(dict(segments=segs, hits=hits))
  if rank == 0:
    print(json.dumps(dict(workload=f'{args.scene}: Lambertian surface source, {n} rays per GPU per step, {args.bins}x{args.bins} fp64 bins '
                                   f'(power-weighted) on the absorber, NCCL all-reduce per step', n_gpus=world, steps=args.steps,
                          rays_per_s=n*world*args.steps/dt, segments_per_s=tot['segments']/dt, binned_hits_per_step=tot['hits']/args.steps,
                          ms_per_step=dt/args.steps*1e3, kernel_ms_per_step=kms/args.steps, allreduce_ms_per_step=ams/args.steps,
                          histogram_sum_last_step=float(total.sum()), histogram_bytes=int(total.size*8))), flush=True)
  if world > 1:
    dist.destroy_process_group()

if __name__ == '__main__':
  main()
