#!/usr/bin/env python
'''
bench.py — traced ray segments/s on the reference's benchmark scenes (BASELINE.json), default = configs[1]:
benchmark/lensesAndMirrors.FCStd, Monte-Carlo mode, 1e8 rays per GPU per step, hit lists stored.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--rays R] [--scene NAME] [--workload hits|binned]
                  [--impl ours|reference]

workload hits   (default) one "step" = one pass of the hot path (odw_trace_mc: sample source -> trace -> append hits)
                over R Monte-Carlo rays per GPU, hit lists kept on the device.  Rays are generated in-kernel from Philox
                counters, so there is no input stream to keep resident; the step's output (hit lists) is far larger than
                L2.  N > 1: one process per GPU (torchrun), disjoint ray ranges per rank, scene replicated, no data-path
                collective -> weak scaling; timing = max over ranks.
workload binned BASELINE.json configs[4]: no hit lists; hits are binned on the device into 1000x1000 fp64 detector
                histograms and the histograms of all ranks are summed with ONE NCCL all-reduce per step INSIDE the timed
                region (the path's only exchange step); e2e = the summed histograms copied to pinned host memory.
scenes          minimal | lensesAndMirrors | lensesAndMirrorsSequential | hugeArray | lambertSource | surfaceSourceTest21
                (read from the reference's own FCStd through the headless importer when a copy travelled with the repo
                under baseline/_ref/scenes/ — __graft_entry__.build() puts it there — else from the exported fixture
                tests/golden/scenes/<scene>.npz; the line says which, and reports scene_export_s / table_build_s).

Prints ONE JSON line (rank 0).  --impl reference times the CPU restatement of the reference's loop (oracle/, all host
threads) — the reference itself needs FreeCAD/OpenCASCADE, absent from this image and from the GPU box
(profiles/r02_freecad_probe.txt).
'''
import argparse
import json
import os
import subprocess
import sys
import tempfile
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
FCSTD_OF = dict(minimal='minimal.FCStd', lensesAndMirrors='lensesAndMirrors.FCStd',
                lensesAndMirrorsSequential='lensesAndMirrorsSequential.FCStd', hugeArray='hugeArray.FCStd',
                lambertSource='lambert-source.FCStd')


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--rays', type=float, default=0, help='Monte-Carlo rays per GPU per step (default 1e8; binned: 1e9)')
  ap.add_argument('--scene', default=None)
  ap.add_argument('--workload', default='hits', choices=['hits', 'binned'])
  ap.add_argument('--bins', type=int, default=1000)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--cpu-sample-rays', type=float, default=0, help='0 = size the CPU sample for ~10 s')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  ap.add_argument('--no-e2e', action='store_true')
  ap.add_argument('--no-plugin', action='store_true', help='skip the e2e_plugin leg (runSimulationIteration + flush)')
  ap.add_argument('--fixture', action='store_true', help='always load the exported scene fixture, never the FCStd')
  args = ap.parse_args()
  if args.scene is None:
    args.scene = 'lambertSource' if args.workload == 'binned' else 'lensesAndMirrors'
  if not args.rays:
    args.rays = 1e9 if args.workload == 'binned' else 1e8
  return args


def workload_string(args):
  'identical in both arms (ours / reference)'
  n = int(args.rays)
  if args.workload == 'binned':
    return (f'{scene_label(args.scene)}, Monte-Carlo (true) mode, {n} rays per GPU per step, no hit lists: detector hits binned on the '
            f'device ({args.bins}x{args.bins} fp64 per detector) and all-reduced over the GPUs every step')
  return f'{scene_label(args.scene)}, Monte-Carlo (true) mode, {n} rays per GPU per step, hit lists stored (RecordHits groups)'


def scene_label(scene):
  if scene == 'lambertSource':
    return 'test/50-old-tests/lambert-source.FCStd'
  if scene == 'surfaceSourceTest21':
    return 'test/21-simulation-modes/main.FCStd'
  return f'benchmark/{scene}.FCStd'


def metric_name(args):
  if args.workload == 'binned':
    return f'traced ray segments/s, {scene_label(args.scene)} with device-binned, all-reduced detector histograms, 1/2/4/8 B200 vs host CPU'
  return METRIC if args.scene == 'lensesAndMirrors' else METRIC.replace('lensesAndMirrors', args.scene)


def load_sim(args):
  '''
  PreparedSimulation + where it came from.  The FCStd (a copy of the reference's own benchmark document under
  baseline/_ref/scenes/, or the reference tree itself where it exists) goes through the headless importer and is timed;
  otherwise the exported fixture is loaded and the line says so.
  '''
  from freecad.optics_design_workbench_b200.simulation.setup import prepare
  info = dict(scene_source='fixture tests/golden/scenes/%s.npz (exported scene; FCStd not present on this machine)' % args.scene,
              scene_export_s=None)
  path = os.path.join(ROOT, 'tests', 'golden', 'scenes', args.scene+'.npz')
  name = FCSTD_OF.get(args.scene)
  if name and not args.fixture:
    for cand in (os.path.join(ROOT, 'baseline', '_ref', 'scenes', name), os.path.join('/root/reference/benchmark', name),
                 os.path.join('/root/reference/test/50-old-tests', name)):
      if os.path.exists(cand):
        path = cand
        info['scene_source'] = 'FCStd through the headless importer (scene_export/fcstd.py): ' + os.path.relpath(cand, ROOT)
        break
  t0 = time.perf_counter()
  sim = prepare(path)
  dt = time.perf_counter()-t0
  if path.lower().endswith('.fcstd'):
    info['scene_export_s'] = dt
  else:
    info['fixture_load_s'] = dt
  t0 = time.perf_counter()
  sim.source_args(0)
  info['table_build_s'] = time.perf_counter()-t0
  return sim, info


def measured_peak():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
  except Exception:
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def ncu_counters(scene):
  'per-launch counters of the dominant kernel from the committed ncu --set full summary (profiles/traffic*.json), or None'
  for name in (f'traffic_{scene}.json', 'traffic.json' if scene == 'lensesAndMirrors' else None):
    if not name:
      continue
    try:
      with open(os.path.join(ROOT, 'profiles', name)) as f:
        return json.load(f)
    except Exception:
      pass
  return None


def fp64_peak():
  'measured DFMA peak of this GPU model (tools/fp64_peak.cu, profiles/r02_fp64_peak.json), else the derived nominal figure'
  try:
    with open(os.path.join(ROOT, 'profiles', 'r02_fp64_peak.json')) as f:
      d = json.load(f)
      return float(d['tflops']), 'measured DFMA stream (profiles/r02_fp64_peak.json)'
  except Exception:
    return 148*64*2*1.965e9/1e12, 'derived: 148 SMs x 64 DFMA/clk x 2 flop x 1.965 GHz'


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


def host_threads():
  'threads the CPU arm may use: the cores this process may run on (torchrun sets OMP_NUM_THREADS=1, which is not a limit of the box)'
  try:
    return max(1, len(os.sched_getaffinity(0)))
  except AttributeError:
    return max(1, os.cpu_count() or 1)


def binning_specs(sim, bins):
  'one power-weighted bins x bins histogram per recording optical group, over the extent of its faces in its dominant plane'
  import numpy as np
  sc = sim.scene
  specs = []
  for g in np.nonzero(sc.groups['record_hits'])[0]:
    f = sc.faces[sc.faces['group'] == g]
    lo, hi = f['aabb_min'].min(axis=0), f['aabb_max'].max(axis=0)
    thin = int(np.argmin(hi-lo))                                  # the detector's normal: its thinnest extent
    ua, va = [a for a in range(3) if a != thin]
    e = np.eye(3)
    specs.append(dict(group=int(g), nu=bins, nv=bins, origin=(0.0, 0.0, 0.0), uaxis=tuple(e[ua]), vaxis=tuple(e[va]),
                      u_range=(float(lo[ua]), float(hi[ua])), v_range=(float(lo[va]), float(hi[va])), weighted=1))
  if not specs:
    raise SystemExit(f'bench.py: scene has no recording optical group to bin')
  return specs


def cpu_trace(orc, sim, args, first, n, threads):
  'the CPU restatement on rays [first, first+n) of the same workload; returns (counts, seconds)'
  sa = sim.source_args(0)
  if args.workload == 'binned':
    cfg = sim.cfg(store_hits=False, binnings=binning_specs(sim, args.bins))
    cap = 16
  else:
    cfg = sim.cfg(store_hits=True)
    cap = max(1024, int(n)*4)
  t0 = time.perf_counter()
  r = orc.trace_mc(sim.scene, sa, cfg, SEED, int(first), int(n), hit_capacity=cap, threads=threads)
  return r['counts'], time.perf_counter()-t0


def cpu_baseline(sim, args, threads, sample_rays, kind_label):
  'oracle (CPU restatement) timed on a bounded sample of the same workload'
  from oracle import Oracle
  orc = Oracle()
  if not sample_rays:
    _, dt = cpu_trace(orc, sim, args, 0, 20000, threads)
    sample_rays = int(min(5e7, max(2e4, 20000/max(dt, 1e-6)*10.0)))
  sample_rays = int(sample_rays)
  counts, dt = cpu_trace(orc, sim, args, 0, sample_rays, threads)
  return dict(value=counts['segments']/dt, unit=UNIT, cores=threads, kind='port',
              sample=f'{sample_rays} MC rays of the same scene/source/seed ({counts["segments"]} segments) in {dt:.2f} s, {kind_label}')


def run_reference(args, rank, world):
  '--impl reference: the CPU restatement on ALL host threads of the box, rank 0 only'
  if rank != 0:
    return
  args.fixture = args.fixture or False
  sim, info = load_sim(args)
  from oracle import Oracle
  orc = Oracle()
  threads = host_threads()
  _, dt = cpu_trace(orc, sim, args, 0, 20000, threads)                                  # size one step for ~3 s of CPU work
  step_rays = int(min(args.rays, max(2e4, 20000/max(dt, 1e-6)*3.0)))
  for w in range(args.warmup):
    cpu_trace(orc, sim, args, w*step_rays, min(step_rays, 20000), threads)
  segs = 0
  t0 = time.perf_counter()
  for k in range(args.steps):
    counts, _ = cpu_trace(orc, sim, args, k*step_rays, step_rays, threads)
    segs += counts['segments']
  dt = time.perf_counter()-t0
  value = segs/dt
  sample = (f'each step = {step_rays} MC rays (bounded sample of the {int(args.rays)}-ray step) on {threads} host threads, '
            f'CPU restatement of the reference loop (oracle/odw_oracle.c, OpenMP over rays); the reference itself '
            f'needs FreeCAD/OpenCASCADE, which neither this image nor the GPU box has (BASELINE.md estimates the real '
            f'reference at 200-500 segments/s)')
  line = dict(impl='reference', metric=metric_name(args), value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
              ms_per_step=dt/args.steps*1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
              data='synthetic',
              config=dict(workload=workload_string(args), rays_per_step_timed=step_rays, host_threads=threads,
                          scene_source=info['scene_source']),
              cpu_baseline=dict(value=value, unit=UNIT, cores=threads, kind='port', sample=sample),
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


def plugin_leg(sim, eng, args, rank, world, barrier, steps):
  '''
  e2e_plugin: what a user of the plugin surface gets — GenericSourceProxy.runSimulationIteration(mode='true', store=...)
  (reference freecad_elements/generic_source.py:51-146) for the step's rays followed by SimulationResults.flush() (pickle
  files of the reference's result tree, results_store.py:369-460), all inside the timed region.
  '''
  import shutil
  from freecad.optics_design_workbench_b200.simulation import results_store, simulation_loop
  from freecad.optics_design_workbench_b200.freecad_elements.generic_source import GenericSourceProxy
  from freecad.optics_design_workbench_b200.freecad_elements import point_source
  base = tempfile.mkdtemp(prefix=f'odw_bench_r{rank}_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
  try:
    store = results_store.SimulationResults(simulationType='true', basePath=base, simulationRunFolder='raw/simulation-run-000000',
                                            flushEverySeconds=1e9, isMaster=(rank == 0))
    ctx = simulation_loop.SimulationContext(sim, eng, seed=SEED, rank=0, world=1)      # this rank traces its own n_rays per step
    src = GenericSourceProxy(ctx, 0)
    per_iter = point_source.rays_per_iteration(sim.source_records[0], sim.settings)
    iterations = max(1, int(args.rays)//per_iter)
    src.runSimulationIteration(mode='true', store=store, iterations=iterations); store.flush()     # warm-up at full size: the page-locked delivery buffers are allocated once, here
    for f in store.writtenFiles:
      os.remove(f)
    barrier()
    t0 = time.perf_counter()
    segs, written = 0, 0
    for k in range(steps):
      n_before = len(store.writtenFiles)               # large batches are written by runSimulationIteration itself (direct writer)
      c = src.runSimulationIteration(mode='true', store=store, iterations=iterations)
      store.flush()
      segs += c['segments']
      for f in store.writtenFiles[n_before:]:
        written += os.path.getsize(f)
        os.remove(f)                                                  # keep the RAM disk bounded; deleting is part of no user's step but cheap
    barrier()
    dt = time.perf_counter()-t0
    return dict(seconds=dt, segments=segs, steps=steps, rays_per_step=iterations*per_iter, file_bytes_per_step=written//max(steps, 1),
                where=base.rsplit('/', 1)[0])
  finally:
    shutil.rmtree(base, ignore_errors=True)


def main():
  args = parse_args()
  claim_stdout()
  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if args.impl == 'reference':
    run_reference(args, rank, world)
    return

  import ctypes as C
  import numpy as np
  import torch
  import torch.distributed as dist
  from freecad.optics_design_workbench_b200 import engine, _abi
  from freecad.optics_design_workbench_b200.simulation import sharding

  if not torch.cuda.is_available():
    raise SystemExit('bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU restatement)')
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  distributed = world > 1
  if distributed:
    dist.init_process_group('nccl', device_id=device)

  n_rays = int(args.rays)
  binned = args.workload == 'binned'
  sim, scene_info = load_sim(args)
  eng = engine.Engine(local_rank)
  dscene = eng.scene(sim.scene)
  sa = sim.source_args(0)
  dsrc = eng.source(sa)
  stream = torch.cuda.ExternalStream(eng.stream_handle(), device=device)

  def barrier():
    if distributed:
      dist.barrier()
    torch.cuda.synchronize()

  if binned:
    specs = binning_specs(sim, args.bins)
    cfg = sim.cfg(store_hits=False, binnings=specs)
    n_bins_total = sum(s['nu']*s['nv'] for s in specs)
    ar0, ar1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_bins = torch.empty(n_bins_total, dtype=torch.float64).pin_memory()

    def step(k, to_host=False):
      'trace + bin on the device, then ONE all-reduce over all detector histograms (contiguous in the engine); returns (counts, kernel ms, all-reduce ms)'
      first = (k*world + rank)*n_rays
      with dscene.trace_mc(dsrc, cfg, SEED, first, n_rays) as res:
        ptr, _ = res.histogram_device(0)                              # binnings are concatenated: one buffer, one collective
        bins = torch.as_tensor(sharding._DevicePointer(ptr, n_bins_total), device=device)
        ar_ms = 0.0
        if distributed:
          ar0.record()
          dist.all_reduce(bins, op=dist.ReduceOp.SUM)                 # NCCL over NVLink, in place, bins never visit the host
          ar1.record()
        if to_host:
          host_bins.copy_(bins, non_blocking=True)
        torch.cuda.current_stream().synchronize()                     # the result (and its bins) is released on exit
        if distributed:
          ar_ms = ar0.elapsed_time(ar1)
        return res.counts, res.kernel_ms, ar_ms
  else:
    cap = int(n_rays*hits_per_ray_bound(sim)) + 1024
    cfg = sim.cfg(store_hits=True, hit_capacity=cap)

    def step(k, to_host=False):
      first = (k*world + rank)*n_rays               # disjoint Philox counter ranges per rank and step
      with dscene.trace_mc(dsrc, cfg, SEED, first, n_rays) as res:
        return res.counts, res.kernel_ms, 0.0

  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()
  # ---- warm-up
  for w in range(args.warmup):
    step(10_000 + w)
  # ---- timed region: device-resident hot path
  barrier()
  t_region0 = time.perf_counter()
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  kernel_ms, segs, hits, launches, allreduce_ms = 0.0, 0, 0, 0, 0.0
  ev0.record(stream)
  for k in range(args.steps):
    c, ms, ar = step(k)
    kernel_ms += ms; allreduce_ms += ar
    segs += c['segments']; hits += c['hits']; launches += c['waves']
    assert c['hits_dropped'] == 0, c
  ev1.record(stream)
  barrier()
  elapsed_ms = ev0.elapsed_time(ev1)
  t_region1 = time.perf_counter()
  clk = clocks.stop(t_region0, t_region1) if rank == 0 else None

  # ---- e2e: the same call through the C ABI with HOST result buffers (pinned), copies inside the timed region
  e2e = None
  if not args.no_e2e:
    if binned:
      step(20_000, to_host=True)
      barrier()
      t0 = time.perf_counter()
      e_segs = 0
      for k in range(args.steps):
        c, _, _ = step(k, to_host=True)
        e_segs += c['segments']
      barrier()
      e2e = dict(seconds=time.perf_counter()-t0, segments=e_segs, d2h=n_bins_total*8, h2d=C.sizeof(_abi.TraceCfg)+3*8+len(specs)*C.sizeof(_abi.Binning),
                 note='odw_trace_mc with device binning + NCCL all-reduce, then the summed histograms copied to pinned host memory every step')
    else:
      # what the plugin's runSimulationIteration requests (freecad_elements/generic_source.py): the four columns of the
      # reference's hit files; the group column only when more than one optical group records hits
      recording = int(np.count_nonzero(sim.scene.groups['record_hits']))
      columns = ('points', 'directions', 'powers', 'is_entering') + (('group',) if recording != 1 else ())
      _arrays, view = eng.pinned_hit_arrays(cap, columns)
      bytes_per_hit = 24+24+8+1+(4 if recording != 1 else 0)
      cfg_host = sim.cfg(store_hits=True, hit_capacity=cap)
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
      e2e = dict(seconds=time.perf_counter()-t0, segments=e_segs, d2h=int(e_hits/args.steps*bytes_per_hit),
                 h2d=C.sizeof(_abi.TraceCfg) + 3*8,     # the call's scalar arguments; MC rays are generated on the device
                 note='odw_trace_mc_host: hit lists (points, directions, powers, isEntering; + group when several groups record) delivered '
                      'into pinned host arrays, device->host copy of chunk c overlapped with the trace of chunk c+1')
      eng.free_pinned(view)
  # ---- e2e_plugin: runSimulationIteration + flush (hit files of the reference's result tree)
  plugin = None
  # (N = 1 only, like cpu_baseline: eight ranks writing 6 GB of hit files each into the same RAM disk measure the box, not the path)
  if not args.no_e2e and not args.no_plugin and not binned and world == 1 and sim.source_records[0].get('proxy') == 'PointSourceProxy':
    plugin = plugin_leg(sim, eng, args, rank, world, barrier, max(1, min(args.steps, 2)))

  # ---- reduce over ranks: max time, summed work
  if distributed:
    t = torch.tensor([elapsed_ms, kernel_ms, e2e['seconds'] if e2e else 0.0, plugin['seconds'] if plugin else 0.0, allreduce_ms],
                     dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    w = torch.tensor([segs, hits, launches, e2e['segments'] if e2e else 0, plugin['segments'] if plugin else 0], dtype=torch.float64, device='cuda')
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    elapsed_ms, kernel_ms_max, e_seconds, p_seconds, allreduce_ms_max = t.tolist()
    segs_all, hits_all, launches_all, e_segs_all, p_segs_all = w.tolist()
  else:
    kernel_ms_max, e_seconds, p_seconds, allreduce_ms_max = kernel_ms, (e2e['seconds'] if e2e else 0.0), (plugin['seconds'] if plugin else 0.0), 0.0
    segs_all, hits_all, launches_all, e_segs_all, p_segs_all = segs, hits, launches, (e2e['segments'] if e2e else 0), (plugin['segments'] if plugin else 0)

  if rank == 0:
    value = segs_all/(elapsed_ms*1e-3)
    peak, peak_src = measured_peak()
    # dominant kernel: algorithmic bytes of this rank's launches / their CUDA-event time on the engine stream
    alg_bytes = segs*BYTES_PER_SEGMENT + hits*BYTES_PER_HIT
    achieved = alg_bytes/(kernel_ms*1e-3)/1e9
    ncu = ncu_counters(args.scene)
    roof = dict(bound='hbm', achieved=achieved, peak=peak, unit='GB/s', frac=achieved/peak, traffic=None, peak_source=peak_src,
                algorithmic_bytes_per_launch=alg_bytes/max(launches, 1), rays_per_launch=n_rays*args.steps/max(launches, 1),
                note='algorithmic bytes = 144 B/segment + 64 B/recorded hit (wavefront formulation, SURVEY.md §8d); the '
                     'register-resident kernel moves far fewer bytes and is bounded by fp64 issue + latency: see roofline.fp64 and dram_frac')
    if ncu and launches:
      scale = (n_rays*args.steps/launches)/ncu.get('rays_per_launch', 2097152)
      roof['traffic'] = ncu['dram_bytes_per_launch']*scale
      roof['dram_frac'] = ncu['dram_bytes_per_launch']/(ncu['duration_ms']*1e-3)/1e9/peak       # actual DRAM bytes of the ncu-captured launch / its duration / peak
      roof['traffic_source'] = f"ncu --set full capture of {ncu.get('kernel')}: {ncu.get('source')} (profiles/); per launch of {ncu.get('rays_per_launch')} rays, scaled to this launch size"
      if ncu.get('fp64_flop_per_launch'):
        pk, pk_src = fp64_peak()
        segs_per_launch_ncu = ncu.get('segments_per_launch') or (segs/(n_rays*args.steps))*ncu.get('rays_per_launch', 2097152)
        flop_per_segment = ncu['fp64_flop_per_launch']/segs_per_launch_ncu
        tfl = flop_per_segment*segs/(kernel_ms*1e-3)/1e12
        roof['fp64'] = dict(pipe_active_pct=ncu.get('fp64_pipe_active_pct'), issue_active_pct=ncu.get('issue_active_pct'),
                            flop_per_segment=flop_per_segment, achieved_tflops=tfl, peak=pk, peak_source=pk_src, frac=tfl/pk,
                            note='flop per segment from the ncu capture (dadd + dmul + 2 dfma thread instructions / segments of that launch); '
                                 'achieved = that x this run\'s segments / kernel time')
    line = dict(
      metric=metric_name(args), value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
      ms_per_step=elapsed_ms/args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
      dtype='f64', data='synthetic',
      config=dict(workload=workload_string(args),
                  rays_per_gpu_per_step=n_rays, seed=hex(SEED), segments_per_ray=segs/(n_rays*args.steps),
                  scene_source=scene_info['scene_source'],
                  l2='no input stream (rays generated in-kernel from Philox counters); ' +
                     (f'each step updates {n_bins_total*8/1e6:.0f} MB of fp64 bins with atomics' if binned else
                      f'each step writes {hits//args.steps*81/1e9:.2f} GB of hit lists, far above the 126 MB L2'),
                  parallelism=(f'rays sharded over {world} GPU(s), scene replicated; one NCCL all-reduce(SUM) of {n_bins_total*8/1e6:.0f} MB of histograms per step'
                               if binned else f'rays sharded over {world} GPU(s), scene replicated, no data-path collective')),
      gpu_launches=int(launches_all),
      kernel_ms_per_step=kernel_ms_max/args.steps,
      rays_per_s=n_rays*world*args.steps/(elapsed_ms*1e-3),
      recorded_hits_per_s=hits_all/(elapsed_ms*1e-3),
      scene_export_s=scene_info.get('scene_export_s'), table_build_s=scene_info.get('table_build_s'),
      roofline=roof, clocks=clk)
    if binned:
      line['allreduce'] = dict(ms_per_step=allreduce_ms_max/args.steps, bytes=n_bins_total*8,
                               kernel='ncclDevKernel_AllReduce_Sum_f64_RING_LL / _TREE_LL (NCCL picks by size; see profiles/ launch list)' if distributed else None,
                               note='CUDA events around dist.all_reduce on the torch stream, max over ranks' if distributed else 'single rank: no collective')
    if e2e:
      line['e2e'] = dict(value=e_segs_all/e_seconds, unit=UNIT, h2d_bytes_per_step=e2e['h2d'], d2h_bytes_per_step=e2e['d2h'], note=e2e['note'])
    if plugin:
      line['e2e_plugin'] = dict(value=p_segs_all/p_seconds, unit=UNIT, steps=plugin['steps'], rays_per_step=plugin['rays_per_step'],
                                file_bytes_per_step=plugin['file_bytes_per_step'],
                                note=f"GenericSourceProxy.runSimulationIteration(mode='true', store=...) + SimulationResults.flush(): hit files of the "
                                     f"reference's result tree written to {plugin['where']} inside the timed region")
    if not args.no_cpu_baseline and world == 1:      # reported at N=1 only (rank 0)
      line['cpu_baseline'] = cpu_baseline(sim, args, 1, args.cpu_sample_rays, 'scalar C restatement (oracle/odw_oracle.c), 1 thread')
    emit(line)
  if distributed:
    dist.barrier()
    dist.destroy_process_group()


def hits_per_ray_bound(sim):
  'rows of hit list to reserve per ray: each recording Absorber stops the ray (1 row); transparent recorders may add rows'
  import numpy as np
  g = sim.scene.groups
  rec = g['record_hits'] != 0
  if not rec.any():
    return 0.05
  if (g['optical_type'][rec] == 3).all():          # only absorbers record: at most one hit per ray
    return 1.05
  return 4.2


if __name__ == '__main__':
  main()
