'''
GPU worker behind the reference's WorkerProcess protocol — the orchestration swap of SURVEY.md §8f-4.

The reference's master (FreeCAD GUI, or FreecadDocument.runSimulation from Jupyter) fans a simulation out over
`WorkerProcessCount` child processes (reference simulation/processes/simulation_loop.py:450-507): each child is started as
`<freecad executable> -c` (worker_process.py:65-69; the executable is taken from $APPIMAGE when that is set, :51-59), gets
Python SOURCE TEXT on its stdin (:139-163: close documents, App.openDocument(<saved FCStd>),
simulation.runSimulation(action=..., slaveInfo=dict(simulationRunFolder=..., parentPid=...))) and is probed for liveness by
asking it to print a random number (:165-184).  Children never talk to each other; they write hit files and progress files
into the master's run folder and stop when the master drops `simulation-is-done` / `simulation-is-canceled`
(freecad_elements/common.py:158-174) or when the parent dies (simulation_loop.py:573-577).

This module is such a child, with a GPU instead of a FreeCAD process:

    APPIMAGE=/path/to/tools/odw-gpu-worker  WorkerProcessCount=<number of GPUs>     (nothing in the reference changes)

It is an interactive Python console on stdin/stdout (what `FreeCAD -c` is) whose namespace holds just enough of `App` and of
`freecad.optics_design_workbench.simulation` for the text the master sends.  runSimulation claims one GPU (lock files, one
worker per device; more workers than devices share them), imports the saved FCStd headless (scene_export/fcstd.py), and
loops: a batch of iterations of every light source on the device -> hit files / progress files in the master's run folder
-> check the flag files and the parent.  Batches grow geometrically (the master, not the worker, evaluates the end criteria
from the progress files, results_store.py:507-512; like in the reference the in-flight batches of all workers overshoot them).
Every worker draws its own 64-bit Philox seed (the reference seeds numpy per worker from pid, time and thread id,
simulation_loop.py:813-820): runs are statistically independent, not reproducible — as in the reference.
'''
import code
import os
import sys
import time
import types


class _Documents:
  'the part of FreeCAD\'s `App` the master\'s text uses: listDocuments / closeDocument / openDocument'
  def __init__(self):
    self.open = {}
    self.ActiveDocument = None

  def listDocuments(self):
    return dict(self.open)

  def closeDocument(self, name):
    self.open.pop(name, None)
    if self.ActiveDocument is not None and self.ActiveDocument.Name == name:
      self.ActiveDocument = None

  def openDocument(self, path):
    name = os.path.splitext(os.path.basename(path))[0]
    doc = types.SimpleNamespace(Name=name, FileName=os.path.realpath(path), getFileName=lambda p=os.path.realpath(path): p)
    self.open[name] = doc
    self.ActiveDocument = doc
    return doc


_GPU_CLAIM = None

def claim_gpu():
  '''
  One worker per device: the first device whose lock file can be locked is this worker's for its lifetime; when every
  device is taken the workers share them by pid.  Returns the device index.
  '''
  global _GPU_CLAIM
  if _GPU_CLAIM is not None:
    return _GPU_CLAIM[0]
  import fcntl, tempfile
  if os.environ.get('ODW_GPU_DEVICE', '') != '':
    _GPU_CLAIM = (int(os.environ['ODW_GPU_DEVICE']), None)
    return _GPU_CLAIM[0]
  from .. import engine as engine_module
  n = engine_module.device_count()
  if n <= 0:
    raise engine_module.EngineError(-2, 'no CUDA device available; the GPU worker has no CPU fallback')
  for i in range(n):
    f = open(os.path.join(tempfile.gettempdir(), f'odw-gpu-claim-{i}.lock'), 'w')
    try:
      fcntl.flock(f, fcntl.LOCK_EX | fcntl.LOCK_NB)
    except OSError:
      f.close()
      continue
    _GPU_CLAIM = (i, f)
    return i
  _GPU_CLAIM = (os.getpid() % n, None)
  return _GPU_CLAIM[0]


class _Simulation:
  'the part of freecad.optics_design_workbench.simulation the master\'s text uses'
  def __init__(self, app):
    self.app = app
    self.isJupyterContext = False
    self._prepared = {}

  def setIsJupyterContext(self, value):
    self.isJupyterContext = bool(value)

  def _prepare(self, path):
    from .setup import prepare
    key = (path, os.path.getmtime(path))
    if key not in self._prepared:
      self._prepared = {key: prepare(path)}
    return self._prepared[key]

  def runSimulation(self, action, slaveInfo={}):
    '''
    simulation.runSimulation(action, slaveInfo) of a WORKER (simulation_loop.py:291, slaveInfo given): results into the
    master's run folder, until the master says stop.  action: 'true' | 'pseudo' (continuous modes are the ones the
    reference fans out, simulation_loop.py:450-477).
    '''
    from . import results_store, simulation_loop
    from .. import engine as engine_module
    from ..freecad_elements.generic_source import GenericSourceProxy
    from ..freecad_elements import point_source
    doc = self.app.ActiveDocument
    if doc is None:
      raise RuntimeError('no document open')
    if action not in ('true', 'pseudo'):
      raise ValueError(f'GPU workers run the continuous simulation modes, not {action!r}')
    if 'simulationRunFolder' not in slaveInfo:
      raise ValueError('slaveInfo without simulationRunFolder: this process only runs as a worker of a master')
    parent = slaveInfo.get('parentPid')
    base = results_store.results_folder_path(doc.FileName)
    sim = self._prepare(doc.FileName)
    eng = engine_module.Engine(claim_gpu())
    seed = int.from_bytes(os.urandom(8), 'little')
    store = results_store.SimulationResults(simulationType=action, basePath=base, simulationRunFolder=slaveInfo['simulationRunFolder'],
                                            isMaster=False)
    ctx = simulation_loop.SimulationContext(sim, eng, seed=seed, rank=0, world=1)
    sources = [GenericSourceProxy(ctx, i) for i in range(len(sim.source_records))]
    per_iter = max(1, sum(point_source.rays_per_iteration(r, sim.settings) for r in sim.source_records))
    max_iterations = max(1, int(os.environ.get('ODW_WORKER_MAX_BATCH_RAYS', 1 << 24))//per_iter)
    iterations = 1
    print(f'odw gpu worker: {eng.device_name()}, {len(sources)} light source(s), run folder {slaveInfo["simulationRunFolder"]}', file=sys.stderr, flush=True)

    def must_stop():
      if simulation_loop.query_status(base, 'simulation-is-done') or simulation_loop.query_status(base, 'simulation-is-canceled'):
        return True
      if parent is not None:
        try:
          os.kill(int(parent), 0)
        except OSError:
          raise RuntimeError(f'parent pid {parent} seems to have died, exiting as well...')   # simulation_loop.py:573-577
      return False

    try:
      while not must_stop():
        ended = False
        for src in sources:
          try:
            src.runSimulationIteration(mode=action, store=store, iterations=iterations)
          except simulation_loop.SimulationEnded:
            ended = True
        store.incrementIterationCount(iterations)
        store.writeDiskIfNeeded()
        store.dumpProgress()                              # the master evaluates the end criteria from these (results_store.py:462-512)
        if ended:
          break
        iterations = min(max_iterations, iterations*4)
    finally:
      store.flush()
      store.dumpProgress()
      ctx.close()
      eng.close()


def console_namespace():
  'namespace of the console + the module entries the master\'s `from freecad.optics_design_workbench import simulation` needs'
  app = _Documents()
  sim = _Simulation(app)
  pkg = types.ModuleType('freecad.optics_design_workbench')
  pkg.__path__ = []
  mod = types.ModuleType('freecad.optics_design_workbench.simulation')
  mod.setIsJupyterContext, mod.runSimulation = sim.setIsJupyterContext, sim.runSimulation
  pkg.simulation = mod
  sys.modules['freecad.optics_design_workbench'] = pkg
  sys.modules['freecad.optics_design_workbench.simulation'] = mod
  import freecad
  freecad.optics_design_workbench = pkg
  return dict(App=app, FreeCAD=app, __name__='__console__')


def main(argv=None):
  argv = sys.argv[1:] if argv is None else argv
  if argv and argv[0] not in ('-c', '--console'):
    print('usage: gpu_worker -c     (an interactive console on stdin, like `FreeCAD -c`)', file=sys.stderr)
    return 2
  console = code.InteractiveConsole(console_namespace())
  sys.ps1 = sys.ps2 = ''
  for raw in sys.stdin:                                   # the master writes '\\r\\n'-terminated lines (worker_process.py:151-163)
    line = raw.rstrip('\r\n')
    try:
      console.push(line)
    except SystemExit:
      break
    except BaseException:                                 # a failing statement is reported and the console lives on, like FreeCAD's
      import traceback
      traceback.print_exc()
    sys.stdout.flush(); sys.stderr.flush()
  return 0


if __name__ == '__main__':
  sys.exit(main())
