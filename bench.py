#!/usr/bin/env python
'''
bench.py — traced ray segments/s on benchmark scene lensesAndMirrors (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--scene NAME] [--impl ours|reference]

One "step" = one pass of the hot path (odw_trace_mc: sample source -> trace -> append hits) over R
Monte-Carlo rays per GPU, hit lists kept on the device.  Rays are generated in-kernel from Philox
counters, so there is no input stream to keep resident; the step's output (hit lists, R*72 B) is far
larger than L2.  N > 1: launched by torchrun, one process per GPU, disjoint ray ranges per rank, scene
replicated, no data-path collective -> weak scaling; timing = max over ranks.

Prints ONE JSON line (rank 0).  --impl reference times the CPU restatement of the reference's loop
(oracle/, all host threads) — the reference itself needs FreeCAD/OpenCASCADE, absent from this image.
'''
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'traced ray segments/s, lensesAndMirrors.FCStd, 1/2/4/8 B200 vs host CPU'
UNIT = 'segments/s'
SEED = 0x0DDB1A5E
BYTES_PER_SEGMENT = 144          # ray state read 72 B + write 72 B
BYTES_PER_HIT = 64               # point 24 + direction 24 + power 8 + isEntering/pad 8


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--rays', type=float, default=1e8, help='Monte-Carlo rays per GPU per step')
  ap.add_argument('--scene', default='lensesAndMirrors')
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--cpu-sample-rays', type=float, default=0, help='0 = size the CPU sample for ~10 s')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  return ap.parse_args()


def load_sim(scene):
  from freecad.optics_design_workbench_b200.simulation.setup import prepare
  return prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', scene+'.npz'))


def measured_peak():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
  except Exception:
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def ncu_traffic():
  'dram bytes per launch of the trace kernel from the committed ncu --set full summary, or None'
  try:
    with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
      return json.load(f)
  except Exception:
    return None


class ClockSampler:
  'samples nvidia-smi SM clocks and throttle reasons while the timed region runs'
  Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index):
    self.gpu = gpu_index
    self.rows = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                    '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.thread = threading.Thread(target=self._read, daemon=True)
      self.thread.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append((time.perf_counter(), line.strip()))

  def stop(self, t0=None, t1=None):
    'summary of the samples taken between perf_counter times t0 and t1 (all samples if too few fall inside)'
    if not self.proc:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
    time.sleep(0.15)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    sm, smax, reasons, power = [], [], set(), []
    inside = [r for t, r in self.rows if t0 is not None and t0 <= t <= t1 + 0.12]
    rows = inside if len(inside) >= 2 else [r for _, r in self.rows]
    for r in rows:
      c = [x.strip() for x in r.split(',')]
      if len(c) < 9:
        continue
      try:
        sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
      except ValueError:
        continue
      for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
        if val.lower().startswith('active'):
          reasons.add(name)
    if not sm:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
    sm.sort()
    return dict(sm_mhz=sm[len(sm)//2], sm_max_mhz=max(smax), power_w_max=max(power), samples=len(sm),
                samples_in_timed_region=len(inside), reasons=sorted(reasons))


def cpu_baseline(sim, threads, sample_rays, kind_label):
  'oracle (CPU restatement) timed on a bounded sample of the same workload'
  from oracle import Oracle
  orc = Oracle()
  sa = sim.source_args(0)
  cfg = sim.cfg(store_hits=True)
  if not sample_rays:
    t0 = time.perf_counter()
    r = orc.trace_mc(sim.scene, sa, cfg, SEED, 0, 20000, hit_capacity=80000, threads=threads)
    dt = time.perf_counter()-t0
    rate = 20000/max(dt, 1e-6)
    sample_rays = int(min(5e7, max(2e4, rate*10.0)))
  sample_rays = int(sample_rays)
  t0 = time.perf_counter()
  r = orc.trace_mc(sim.scene, sa, cfg, SEED, 0, sample_rays, hit_capacity=sample_rays*4, threads=threads)
  dt = time.perf_counter()-t0
  used = orc.max_threads() if threads == 0 else threads
  return dict(value=r['counts']['segments']/dt, unit=UNIT, cores=used, kind='port',
              sample=f'{sample_rays} MC rays of the same scene/source/seed ({r["counts"]["segments"]} segments) in {dt:.2f} s, '
                     f'{kind_label}'), r['counts'], dt


def run_reference(args, rank, world):
  if rank != 0:
    return
  sim = load_sim(args.scene)
  from oracle import Oracle
  threads = 0
  orc = Oracle()
  ncores = orc.max_threads()
  sa = sim.source_args(0)
  cfg = sim.cfg(store_hits=True)
  # size one step for ~3 s of CPU work
  t0 = time.perf_counter()
  orc.trace_mc(sim.scene, sa, cfg, SEED, 0, 20000, hit_capacity=80000, threads=threads)
  rate = 20000/max(time.perf_counter()-t0, 1e-6)
  step_rays = int(min(args.rays, max(2e4, rate*3.0)))
  for w in range(args.warmup):
    orc.trace_mc(sim.scene, sa, cfg, SEED, w*step_rays, min(step_rays, 20000), hit_capacity=80000, threads=threads)
  segs = 0
  t0 = time.perf_counter()
  for k in range(args.steps):
    r = orc.trace_mc(sim.scene, sa, cfg, SEED, k*step_rays, step_rays, hit_capacity=step_rays*4, threads=threads)
    segs += r['counts']['segments']
  dt = time.perf_counter()-t0
  value = segs/dt
  sample = (f'each step = {step_rays} MC rays of {args.scene} (bounded sample of the {int(args.rays)}-ray step), '
            f'CPU restatement of the reference loop (oracle/odw_oracle.c, OpenMP over rays); the reference itself '
            f'needs FreeCAD/OpenCASCADE which this image does not have')
  line = dict(impl='reference', metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
              ms_per_step=dt/args.steps*1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
              data='synthetic',
              config=dict(workload=f'benchmark/{args.scene}.FCStd, Monte-Carlo (true) mode, {int(args.rays)} rays per GPU per step',
                          rays_per_step_timed=step_rays),
              cpu_baseline=dict(value=value, unit=UNIT, cores=ncores, kind='port', sample=sample),
              e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
              gpu_launches=0)
  emit(line)


_RESULT_OUT = None

def claim_stdout():
  '''
  The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout when the
  image sets NCCL_DEBUG=VERSION), so file descriptor 1 is pointed at stderr for the whole run and the result line goes
  to a private duplicate of the original stdout.
  '''
  global _RESULT_OUT
  if _RESULT_OUT is None:
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)

def emit(line):
  print(json.dumps(line), file=_RESULT_OUT or sys.stdout, flush=True)


def main():
  args = parse_args()
  claim_stdout()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if args.impl == 'reference':
    run_reference(args, rank, world)
    return

  import numpy as np
  import torch
  import torch.distributed as dist
  from freecad.optics_design_workbench_b200 import engine, _abi

  if not torch.cuda.is_available():
    raise SystemExit('bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU restatement)')
  torch.cuda.set_device(local_rank)
  distributed = world > 1
  if distributed:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

  n_rays = int(args.rays)
  sim = load_sim(args.scene)
  eng = engine.Engine(local_rank)
  dscene = eng.scene(sim.scene)
  sa = sim.source_args(0)
  dsrc = eng.source(sa)
  cap = int(n_rays*1.05) + 1024
  cfg = sim.cfg(store_hits=True, hit_capacity=cap)
  stream = torch.cuda.ExternalStream(eng.stream_handle(), device=torch.device('cuda', local_rank))

  def step(k):
    first = (k*world + rank)*n_rays               # disjoint Philox counter ranges per rank and step
    return dscene.trace_mc(dsrc, cfg, SEED, first, n_rays)

  def barrier():
    if distributed:
      dist.barrier()
    torch.cuda.synchronize()

  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()
  # ---- warm-up
  for w in range(args.warmup):
    step(10_000 + w).close()
  # ---- timed region: device-resident hot path
  barrier()
  t_region0 = time.perf_counter()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  kernel_ms, segs, hits, launches = 0.0, 0, 0, 0
  ev0.record(stream)
  for k in range(args.steps):
    with step(k) as res:
      c = res.counts
      kernel_ms += res.kernel_ms
      segs += c['segments']; hits += c['hits']; launches += c['waves']
      assert c['hits_dropped'] == 0, c
  ev1.record(stream)
  barrier()
  elapsed_ms = ev0.elapsed_time(ev1)
  t_region1 = time.perf_counter()
  sm_clock_in_kernel = None

  # ---- e2e: the same call with HOST result buffers (pinned), D2H inside the timed region
  e2e = None
  if not args.no_e2e:
    # what the plugin's runSimulationIteration requests (freecad_elements/generic_source.py): the four columns of the
    # reference's hit files; the group column only when more than one optical group records hits
    recording = int(np.count_nonzero(sim.scene.groups['record_hits']))
    columns = ('points', 'directions', 'powers', 'is_entering') + (('group',) if recording != 1 else ())
    _arrays, view = eng.pinned_hit_arrays(cap, columns)
    bytes_per_hit = 24+24+8+1+(4 if recording != 1 else 0)
    import ctypes as C
    cfg_host = sim.cfg(store_hits=True)
    def e2e_step(k):
      first = (k*world + rank)*n_rays
      return dscene.trace_mc_host(dsrc, cfg_host, SEED, first, n_rays, view)
    e2e_step(20_000)
    barrier()
    t0 = time.perf_counter()
    e_segs, e_hits = 0, 0
    for k in range(args.steps):
      c, got = e2e_step(k)
      e_segs += c['segments']; e_hits += got
      assert c['hits_dropped'] == 0, c
    barrier()
    e_dt = time.perf_counter()-t0
    d2h = int(e_hits/args.steps*bytes_per_hit)
    h2d = C.sizeof(_abi.TraceCfg) + 3*8     # the call's scalar arguments; MC rays are generated on the device
    e2e = dict(seconds=e_dt, segments=e_segs, d2h=d2h, h2d=h2d)
  clk = clocks.stop(t_region0, t_region1) if rank == 0 else None

  # ---- reduce over ranks: max time, summed work
  if distributed:
    t = torch.tensor([elapsed_ms, kernel_ms, e2e['seconds'] if e2e else 0.0], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    w = torch.tensor([segs, hits, launches, e2e['segments'] if e2e else 0], dtype=torch.float64, device='cuda')
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    elapsed_ms, kernel_ms_max, e_seconds = t.tolist()
    segs_all, hits_all, launches_all, e_segs_all = w.tolist()
  else:
    kernel_ms_max, e_seconds = kernel_ms, (e2e['seconds'] if e2e else 0.0)
    segs_all, hits_all, launches_all, e_segs_all = segs, hits, launches, (e2e['segments'] if e2e else 0)

  if rank == 0 and clk is None:
    clk = clocks.stop(t_region0, t_region1)
  if rank == 0:
    value = segs_all/(elapsed_ms*1e-3)
    peak, peak_src = measured_peak()
    # dominant kernel = trace_kernel; algorithmic bytes per launch / its mean CUDA-event duration (this rank)
    alg_bytes = segs*BYTES_PER_SEGMENT + hits*BYTES_PER_HIT
    achieved = alg_bytes/(kernel_ms*1e-3)/1e9
    traffic = ncu_traffic() if args.scene == 'lensesAndMirrors' else None     # the committed ncu capture is of this scene's kernel
    line = dict(
      metric=METRIC if args.scene == 'lensesAndMirrors' else METRIC.replace('lensesAndMirrors', args.scene), value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
      ms_per_step=elapsed_ms/args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
      dtype='f64', data='synthetic',
      config=dict(workload=f'benchmark/{args.scene}.FCStd, Monte-Carlo (true) mode, {n_rays} rays per GPU per step, '
                           f'hit lists stored (RecordHits groups)',
                  rays_per_gpu_per_step=n_rays, seed=hex(SEED), segments_per_ray=segs/(n_rays*args.steps),
                  l2='no input stream (rays generated in-kernel from Philox counters); each step writes '
                     f'{hits//args.steps*72/1e9:.2f} GB of hit lists, far above the 126 MB L2',
                  parallelism=f'rays sharded over {world} GPU(s), scene replicated, no data-path collective'),
      gpu_launches=int(launches_all),
      kernel_ms_per_step=kernel_ms_max/args.steps,
      rays_per_s=n_rays*world*args.steps/(elapsed_ms*1e-3),
      recorded_hits_per_s=hits_all/(elapsed_ms*1e-3),
      roofline=dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved/peak,
                    traffic=(traffic['dram_bytes_per_launch']*(n_rays*args.steps/launches)/traffic.get('rays_per_launch', 2097152)
                             if traffic and launches else None),
                    peak_source=peak_src,
                    algorithmic_bytes_per_launch=alg_bytes/max(launches, 1), rays_per_launch=n_rays*args.steps/max(launches, 1),
                    launch_overlap='the launches of a step run on 4 streams; achieved = bytes of all launches / CUDA-event time from the first '
                                   'launch to the last completion (per-launch durations overlap); traffic = ncu dram bytes of a 2^21-ray launch '
                                   'scaled to this launch size (profiles/traffic.json)',
                    note='algorithmic bytes = 144 B/segment + 64 B/recorded hit (wavefront formulation, SURVEY.md §8d); the '
                         'register-resident kernel moves far fewer bytes and is bounded by fp64 issue, see DESIGN.md'),
      clocks=clk)
    if e2e:
      line['e2e'] = dict(value=e_segs_all/e_seconds, unit=UNIT, h2d_bytes_per_step=e2e['h2d'], d2h_bytes_per_step=e2e['d2h'],
                         note='odw_trace_mc_host: hit lists (points, directions, powers, isEntering; + group when several groups record) delivered into pinned host arrays, '
                              'device->host copy of chunk c overlapped with the trace of chunk c+1')
    if not args.no_cpu_baseline and world == 1:      # reported at N=1 only (rank 0)
      cb, _, _ = cpu_baseline(sim, 1, args.cpu_sample_rays, 'scalar C restatement (oracle/odw_oracle.c), 1 thread')
      line['cpu_baseline'] = cb
    emit(line)
  if distributed:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
