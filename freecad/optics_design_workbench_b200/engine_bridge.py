'''
Bridge for the reference workbench: the two functions INTEGRATION.md's stub calls from
GenericSourceProxy.runSimulationIteration (reference freecad_elements/generic_source.py:51-146).

  available()                      -> bool: library built and a CUDA device present
  run_iteration(proxy, obj, ...)   -> traces one iteration of light source `obj` on the GPU and records the
                                      hits in the run folder of the reference's `store`

`obj` is the FreeCAD document object of the light source.  The scene is imported headless from the SAVED
project file (`obj.Document.FileName`; the reference saves the document before it starts its workers,
simulation_loop.py:464-477), once per file version.  Hits are written by our writer into the run folder of the
reference's SimulationResults (`store.basePath`, `store.simulationRunFolder`) in the reference's own file
format, and the reference store's progress counters are advanced, so end criteria, the progress window and the
loaders behave as before.
'''

import os

from .simulation import results_store, simulation_loop
from .simulation.setup import prepare

_engine_factory = None
_available = None
_contexts = {}          # (path, mtime) -> (SimulationContext, {source name: GenericSourceProxy})
_writers = {}           # run folder -> results_store.SimulationResults


def set_engine_factory(factory):
  'tests inject an engine here; the product default is engine.Engine(LOCAL_RANK)'
  global _engine_factory, _available
  _engine_factory, _available = factory, None
  _contexts.clear()


def _make_engine():
  if _engine_factory is not None:
    return _engine_factory()
  from . import engine
  return engine.Engine(int(os.environ.get('LOCAL_RANK', '0')))


def available():
  global _available
  if _available is None:
    try:
      _make_engine().close()
      _available = True
    except Exception:
      _available = False
  return _available


def _context_for(path):
  key = (os.path.realpath(path), os.path.getmtime(path))
  if key not in _contexts:
    for old in [k for k in _contexts if k[0] == key[0]]:
      _contexts.pop(old)[0].close()
    sim = prepare(path)
    ctx = simulation_loop.SimulationContext(sim, _make_engine())
    from .freecad_elements.generic_source import GenericSourceProxy
    proxies = {r['name']: GenericSourceProxy(ctx, i) for i, r in enumerate(sim.source_records)}
    _contexts[key] = (ctx, proxies)
  return _contexts[key]


def _writer_for(store):
  key = f'{store.basePath}/{store.simulationRunFolder}'
  if key not in _writers:
    _writers[key] = results_store.SimulationResults(getattr(store, 'simulationType', 'true'), store.basePath,
                                                    simulationRunFolder=store.simulationRunFolder, isMaster=False,
                                                    flushEverySeconds=getattr(store, 'flushEverySeconds', 5))
  return _writers[key]


def run_iteration(proxy, obj, *, mode, store=False, useInitialConditions=None, iterations=1, **kwargs):
  path = obj.Document.FileName
  ctx, proxies = _context_for(path)
  if obj.Name not in proxies:
    raise RuntimeError(f'light source {obj.Name} not found in the saved project {path}; save the document first')
  writer = _writer_for(store) if store else False
  before = (writer.totalTracedRays, writer.progressDict()['totalRecordedHits']) if writer else (0, 0)
  counts = proxies[obj.Name].runSimulationIteration(mode=mode, store=writer, useInitialConditions=useInitialConditions,
                                                    iterations=iterations, **kwargs)
  if writer:
    writer.writeDiskIfNeeded()
    # advance the reference store's own counters (results_store.py:340-346,472-476)
    store.totalTracedRays += writer.totalTracedRays-before[0]
    store.totalRecordedHits += writer.progressDict()['totalRecordedHits']-before[1]
  return counts


def flush(store):
  'call from the reference store\'s flush() / at simulation end so buffered GPU hits reach the disk'
  key = f'{store.basePath}/{store.simulationRunFolder}'
  if key in _writers:
    _writers[key].flush()
