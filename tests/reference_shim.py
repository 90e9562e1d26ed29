'''
Imports pieces of the REFERENCE (read-only tree at /root/reference) for tests marked `reference`:
its sampler, its result-store helpers and its jupyter_utils loaders.  The package __init__ files are
bypassed with synthetic parent packages and the absent third-party modules (matplotlib, seaborn,
atomicwrites) are stubbed — none of them is used by the functions the tests call.  Nothing is copied.
'''
import contextlib
import os
import sys
import types

ROOT = '/root/reference/freecad/optics_design_workbench'


def _stub(name, **attrs):
  m = sys.modules.get(name)
  if m is None:
    m = types.ModuleType(name)
    sys.modules[name] = m
  for k, v in attrs.items():
    setattr(m, k, v)
  return m


def load():
  'returns a namespace with .distributions, .results_store, .io, .hits, .histogram, .RawFolder (None if unavailable)'
  if 'odw_ref' in sys.modules and hasattr(sys.modules['odw_ref'], '_shim'):
    return sys.modules['odw_ref']._shim
  try:
    import matplotlib.pyplot  # noqa: F401
  except ImportError:
    plt = _stub('matplotlib.pyplot', __all__=[])
    mpl = _stub('matplotlib', pyplot=plt, ticker=_stub('matplotlib.ticker'))
    mpl.__path__ = []
  try:
    import seaborn  # noqa: F401
  except ImportError:
    _stub('seaborn')
  try:
    import atomicwrites  # noqa: F401
  except ImportError:
    @contextlib.contextmanager
    def atomic_write(path, mode='w', overwrite=False):
      with open(path, mode) as f:
        yield f
    _stub('atomicwrites', atomic_write=atomic_write)
  pkg = _stub('odw_ref'); pkg.__path__ = [ROOT]
  sim = _stub('odw_ref.simulation'); sim.__path__ = [ROOT+'/simulation']
  proc = _stub('odw_ref.simulation.processes', isMasterProcess=lambda: None, simulatingDocument=lambda: None)
  sim.processes = proc
  ju = _stub('odw_ref.jupyter_utils'); ju.__path__ = [ROOT+'/jupyter_utils']
  from odw_ref import io, distributions
  from odw_ref.simulation import results_store
  sim.findPathsAndSanitize = results_store.findPathsAndSanitize
  sim.updateResultEntry = results_store.updateResultEntry
  sim.results_store = results_store
  from odw_ref.jupyter_utils import histogram, hits
  raw_folder = None
  try:
    from odw_ref.jupyter_utils import freecad_document
    raw_folder = freecad_document.RawFolder
  except Exception:
    pass
  ns = types.SimpleNamespace(io=io, distributions=distributions, results_store=results_store, hits=hits,
                             histogram=histogram, RawFolder=raw_folder)
  pkg._shim = ns
  return ns
