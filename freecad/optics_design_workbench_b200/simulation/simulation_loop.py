'''
Simulation driver around the engine: runSimulation(action) with the reference's action names and
end criteria, writing the reference's result tree.

Mirrors reference simulation/processes/simulation_loop.py:291-775 for the part that surrounds the hot
path ("mainloop A", :544-632: for every light source runSimulationIteration, count the iteration, write
to disk when due, stop when the aggregated progress exceeds EndAfter*).  The reference's worker-process
fan-out (:450-507) is replaced by GPU ranks (simulation/sharding.py); its GUI / flag-file state machine
(:174-269) is out of scope and stays in the reference.
'''

import os
import time

import numpy as np

from . import results_store, sharding
from .setup import prepare, PreparedSimulation
from ..freecad_elements.generic_source import GenericSourceProxy

class SimulationEnded(RuntimeError):
  'control flow like the reference (freecad_elements/common.py:155): a light source ran dry / the run was finished'


# ---- flag files (reference simulation/processes/simulation_loop.py:174-269) ------------------------------------
# Empty files in <doc>.OpticsDesign/ are the reference's control channel: the Qt progress window, the toolbar's stop
# button and FreecadDocument.runSimulation(endIf=...) (jupyter_utils/freecad_document.py:711-746) drop
# `simulation-is-canceled` / `simulation-is-done` there, and everybody reads `simulation-is-running`.
def _status_path(base, name):
  return f'{base}/{name}'

def query_status(base, name):
  return os.path.exists(_status_path(base, name))

def set_status(base, name, state):
  path = _status_path(base, name)
  if state and not os.path.exists(path):
    os.makedirs(base, exist_ok=True)
    with open(path, 'w'):
      pass
  elif not state and os.path.exists(path):
    try:
      os.remove(path)
    except FileNotFoundError:
      pass

def cancelSimulation(basePath):
  'what the stop action does (simulation_loop.py:249-251)'
  if query_status(basePath, 'simulation-is-running'):
    set_status(basePath, 'simulation-is-canceled', True)


DEFAULT_SEED = 0x0DDB1A5E
ACTIONS = ('fans', 'singlepseudo', 'singletrue', 'pseudo', 'true')


class SimulationContext:
  '''
  Everything one simulation run shares: the prepared project, the engine objects on this rank's GPU, the
  Philox seed and the allocator of global ray indices.  Replaces the reference's per-segment look-ups of
  the active settings and the document (ray.py:46,286; find.py:79-141).
  '''
  def __init__(self, sim, engine, seed=DEFAULT_SEED, rank=0, world=1):
    self.sim, self.engine, self.seed = sim, engine, int(seed)
    self.rank, self.world = int(rank), int(world)
    self.device_scene = engine.scene(sim.scene)
    self._device_sources = {}
    self._next_ray = {}                   # per light source: next unused global ray index
    self._replay = {}                     # per replay source: its stock of rays
    self._pinned = {}                     # page-locked hit arrays by column set

  def device_source(self, index):
    if index not in self._device_sources:
      self._device_sources[index] = self.engine.source(self.sim.source_args(index))
    return self._device_sources[index]

  def cfg(self, source_record, **overrides):
    'odw_trace_cfg of the active settings; the per-source scales are applied inside the engine for MC sources'
    return self.sim.cfg(**overrides)

  def claim_rays(self, source_index, n_global):
    'reserve the next n_global ray indices of a source for ALL ranks and return this rank\'s shard (first, n)'
    first = self._next_ray.get(source_index, 0)
    self._next_ray[source_index] = first+int(n_global)
    return sharding.shard_range(first, n_global, self.rank, self.world)

  def pinned_hits(self, capacity, columns):
    'page-locked hit arrays of at least `capacity` rows with these columns, reused from call to call'
    have = self._pinned.get(columns)
    if have is None or have[2] < capacity:
      if have is not None and hasattr(self.engine, 'free_pinned'):
        self.engine.free_pinned(have[1])                 # the superseded, smaller buffers go back to the driver
        del self._pinned[columns]
      if hasattr(self.engine, 'pinned_hit_arrays'):
        arrays, view = self.engine.pinned_hit_arrays(capacity, columns)
      else:                                              # engines without page-locked memory (the CPU test double)
        from .. import _abi
        h = _abi.HitArrays(capacity)
        arrays, view = {k: getattr(h, k) for k in columns}, h.view
        for k in ('points', 'directions', 'powers', 'is_entering', 'ray_index', 'group', 'bounce', 'face_id', 'medium'):
          if k not in columns:
            setattr(view, k, None)
        self._keep = h
      have = self._pinned[columns] = (arrays, view, capacity)
    return have[0], have[1]

  def replay_stock(self, index, loader):
    if index not in self._replay:
      self._replay[index] = loader()
    return self._replay[index]

  def close(self):
    for s in self._device_sources.values():
      s.close()
    self.device_scene.close()


def collect_global_info(sim):
  'global-info.pkl (reference freecad_elements/__init__.py:48-116), from what the headless importer knows'
  eye = np.eye(4)
  def entry(name, label, props, gpM):
    gpM = np.asarray(gpM, dtype=np.float64).reshape(4, 4)
    return dict(name=name, label=label, properties=props,
                placementPathsAndMatrices=[dict(path=name, gpM=gpM, gpMi=np.linalg.inv(gpM), pM=eye, pMi=eye)])
  sources = [entry(r['name'], r['label'], {k: v for k, v in r.items() if k[:1].isupper()}, r['gpM'])
             for r in sim.source_records]
  scene = sim.scene
  from ..scene_export.scene import OPTICAL_TYPES
  objects = [entry(n, l, dict(OpticalType=OPTICAL_TYPES[int(g['optical_type'])], RefractiveIndex=float(g['refractive_index']),
                              Reflectivity=float(g['reflectivity']), RecordHits=bool(g['record_hits'])), eye)
             for n, l, g in zip(scene.group_names, scene.group_labels, scene.groups)]
  return dict(activeSimulationSettings={k: v for k, v in sim.settings.items()}, lightSources=sources, opticalObjects=objects)


def _iterations_until_end(store, rays_per_iteration_all_sources, max_batch_rays):
  '''
  How many iterations to put into the next engine call.  A reference worker checks the end criteria after every
  iteration (simulation_loop.py:601-629) and stops at the first count STRICTLY above the limit; with rays and
  iterations that point is known in advance, with hits it is not (batches grow while the hit rate is measured).
  '''
  n = max(1, int(max_batch_rays//max(1, rays_per_iteration_all_sources)))
  if np.isfinite(store.endAfterIterations):
    n = min(n, int(store.endAfterIterations)+1-store.totalIterations)
  if np.isfinite(store.endAfterRays):
    left = store.endAfterRays-store.totalTracedRays
    n = min(n, int(left//rays_per_iteration_all_sources)+1)
  return max(1, n)


def runSimulation(project, action, *, engine=None, basePath=None, seed=DEFAULT_SEED, settings=None,
                  maxBatchRays=1 << 24, flushEverySeconds=5, keepProgressFiles=False):
  '''
  project   path of a .FCStd (or scene fixture .npz) or a PreparedSimulation
  action    'fans' | 'singletrue' | 'true'   ('pseudo', 'singlepseudo': not on the engine yet)
  engine    engine.Engine of this rank's GPU (default: Engine(LOCAL_RANK)); injected by the CPU tests
  settings  overrides of the active OpticalSimulationSettings (EndAfterRays=…, RaysPerIteration=…)
  Returns the run folder path (what FreecadDocument.runSimulation wraps into a RawFolder).

  Under torch.distributed (one process per GPU) every rank traces its shard of each batch and writes its own hit
  files into the SAME run folder; rank 0 creates the folder, writes global-info.pkl and the master progress file;
  progress counters are all-reduced after every batch to evaluate the end criteria.
  '''
  if action not in ACTIONS:
    raise ValueError(f'unknown simulation action {action}')
  sim = project if isinstance(project, PreparedSimulation) else prepare(project)
  if settings:
    sim.settings.update(settings)
  rank, world = sharding.rank_and_world()
  if engine is None:
    from .. import engine as engine_module
    engine = engine_module.Engine(int(os.environ.get('LOCAL_RANK', '0')))
  if basePath is None:
    if isinstance(project, str) and project.lower().endswith('.fcstd'):
      basePath = results_store.results_folder_path(project)
    else:
      raise ValueError('basePath (the <name>.OpticsDesign folder) is required when the project is not a .FCStd path')
  continuous = action in ('true', 'pseudo')
  mode = action
  if rank == 0:                                    # simulation_loop.py:323-334: fresh flags for this run
    set_status(basePath, 'simulation-is-canceled', False)
    set_status(basePath, 'simulation-is-done', False)
    set_status(basePath, 'simulation-is-running', True)
  s = sim.settings
  run_folder = results_store.generate_simulation_folder_name(basePath) if rank == 0 else None
  if rank == 0:
    os.makedirs(f'{basePath}/{run_folder}', exist_ok=True)
  run_folder = sharding.broadcast_object(run_folder)
  store = results_store.SimulationResults(
    simulationType=action, basePath=basePath, simulationRunFolder=run_folder, flushEverySeconds=flushEverySeconds,
    endAfterIterations=s.get('EndAfterIterations', np.inf) if continuous else np.inf,
    endAfterRays=s.get('EndAfterRays', np.inf) if continuous else np.inf,
    endAfterHits=s.get('EndAfterHits', np.inf) if continuous else np.inf, isMaster=(rank == 0))
  if rank == 0:
    store.dumpGlobalInfo(collect_global_info(sim))
  ctx = SimulationContext(sim, engine, seed=seed, rank=rank, world=world)
  sources = [GenericSourceProxy(ctx, i) for i in range(len(sim.source_records))]
  run_error, clean_exit = None, True
  try:
    if not continuous:
      # one iteration (simulation_loop.py:342-411: single shots and fans are not continuous)
      for src in sources:
        if src.record.get('proxy') == 'ReplaySourceProxy' and mode != 'fans':
          try:
            src.runSimulationIteration(mode='true', store=store, iterations=1)
          except SimulationEnded:
            pass
          continue
        if mode == 'fans':
          # fans are a short deterministic list: rank 0 traces them (the reference's multicorefans mailbox,
          # results_store.py:679-738, distributes chunks of the same list; not worth it for <= 1e3 rays)
          if rank == 0:
            src.runSimulationIteration(mode='fans', store=store)
        else:
          src.runSimulationIteration(mode='pseudo' if mode == 'singlepseudo' else 'true', store=store, iterations=1)
      store.incrementIterationCount()
    else:
      has_replay = any(r.get('proxy') == 'ReplaySourceProxy' for r in sim.source_records)
      if not has_replay and not any(np.isfinite(v) for v in (store.endAfterIterations, store.endAfterRays, store.endAfterHits)):
        raise ValueError("continuous simulation without any end criterion (EndAfterRays/Hits/Iterations all 'inf')")
      from ..freecad_elements import point_source
      per_iter = sum(point_source.rays_per_iteration(r, s) for r in sim.source_records)
      batch_rays = min(maxBatchRays, max(per_iter, 1 << 16))
      while True:
        k = _iterations_until_end(store_global(store, world), per_iter, batch_rays)
        ended, failure = False, None
        try:
          for src in sources:
            try:
              src.runSimulationIteration(mode='pseudo' if mode == 'pseudo' else 'true', store=store, iterations=k)
            except SimulationEnded:
              ended = True                                        # a replay source ran out of rays (replay_source.py:160-161)
          store.incrementIterationCount(k)
          store.writeDiskIfNeeded()
        except Exception as exc:                                  # reported to the other ranks below, raised after the loop
          failure = exc
        # ONE collective per batch carries the counters and every reason to leave the loop, so that all ranks take the same
        # branch: a rank whose replay source ran dry, a rank that failed and a flag file seen by one rank only stop everybody
        stop = query_status(basePath, 'simulation-is-canceled') or query_status(basePath, 'simulation-is-done')
        total = sharding.all_reduce_counters(dict(totalTracedRays=store.totalTracedRays,
                                                   totalRecordedHits=store.progressDict()['totalRecordedHits'],
                                                   _stop=int(stop), _ended=int(ended), _failed=int(failure is not None)))
        stop, ended, failed = bool(total.pop('_stop')), bool(total.pop('_ended')), bool(total.pop('_failed'))
        if failed:
          run_error = failure if failure is not None else RuntimeError('another rank failed during the simulation; see its log')
          break
        total['totalIterations'] = store.totalIterations        # iterations are global (every rank takes part in each)
        total['totalRecordedRays'] = 0
        store._global = total
        if rank == 0:
          store.dumpMasterProgress(total)
        if store.isEndReached(total):
          if rank == 0:
            set_status(basePath, 'simulation-is-done', True)      # results_store.py:507-512 -> setIsFinished(True)
          break
        if ended or stop:
          break
        batch_rays = min(maxBatchRays, batch_rays*4)            # grow while only EndAfterHits is pending
  except BaseException:
    clean_exit = False                                            # an exception outside the batch loop: do not wait for ranks that may never arrive
    raise
  finally:
    store.flush()
    if clean_exit:
      sharding.barrier()
    if rank == 0:
      set_status(basePath, 'simulation-is-running', False)        # simulation_loop.py:726-775
    if not keepProgressFiles:
      if rank == 0:
        store.cleanup()
      else:
        store._cleanedUp = True
    ctx.close()
  if run_error is not None:
    raise run_error
  return store.runFolderPath()


class _GlobalView:
  'the store\'s limits with the all-reduced (global) counters, for _iterations_until_end'
  def __init__(self, store, total):
    self.endAfterIterations, self.endAfterRays = store.endAfterIterations, store.endAfterRays
    self.totalIterations = total.get('totalIterations', 0)
    self.totalTracedRays = total.get('totalTracedRays', 0)


def store_global(store, world):
  return _GlobalView(store, getattr(store, '_global', dict(totalIterations=0, totalTracedRays=0)))
