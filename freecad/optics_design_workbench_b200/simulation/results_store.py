'''
Hit writer: keeps the reference's on-disk result tree so RawFolder.loadHits / Hits / the progress
window read GPU results unchanged.

Mirrors SimulationResults of the reference (reference simulation/results_store.py):
  folder naming            :49-72     raw/simulation-run-%06d, uid-<uuid4> marker :294-296
  _makeFilename            :352-367   source-<Label>/object-<Label>/<ms>-pid<PID>-thread<TID>-<kind>.pkl
  flush                    :369-460   one pickle per (source, object): dict(source, obj, points, directions,
                                      powers, isEntering [+ StoreHit* metadata keys])
  dumpProgress/getProgress :462-540   progress/<fingerprint>-<hex>.pkl, progress/master-%09d, end criteria (strict >)
  addRayHit                :641-648
What differs: hits arrive as ARRAYS (one batch per engine call) instead of one Python list entry per
hit, so flush() concatenates arrays instead of walking a list.  The files are the same.
'''

import os
import pickle
import random
import threading
import time
import uuid

import numpy as np

_README_TEXT = '''
This folder was created by the Optics Design Workbench simulation engine. The raw subfolder contains one
simulation-run-XXXXXX folder per simulation run with the recorded hits (and rays); the notebooks subfolder
is the place for jupyter notebooks that analyse them with freecad.optics_design_workbench.jupyter_utils.
'''.strip()

# metadata keys the reference can attach to every hit (ray.py:57-65 filters ray metadata by the StoreHit*
# switches of the settings object), in their original camelCase
HIT_METADATA_KEYS = ('initPoint', 'initDirection', 'initPower', 'initWavelength', 'initPhi', 'initTheta',
                     'rayIndex', 'fanIndex', 'totalFanCount', 'totalRaysInFan')


def atomic_write_bytes(path, data, overwrite=True):
  'write-then-rename (the reference uses the atomicwrites package for the same purpose)'
  if not overwrite and os.path.exists(path):
    raise FileExistsError(path)
  tmp = f'{path}.tmp{os.getpid()}-{threading.get_ident()}'
  with open(tmp, 'wb') as f:
    f.write(data)
  os.replace(tmp, path)


def results_folder_path(fcstd_path):
  'results_store.py:184-200: <dir>/<name>.OpticsDesign next to <name>.FCStd'
  base, fname = os.path.split(os.path.realpath(fcstd_path))
  if fname.lower().endswith('.fcstd'):
    fname = fname[:-6]
  return f'{base}/{fname}.OpticsDesign'


def latest_run_index(base_path):
  folder = base_path+'/raw'
  if os.path.exists(folder):
    idx = [int(f[len('simulation-run-'):]) if f[len('simulation-run-'):].isnumeric() else -1
           for f in os.listdir(folder) if f.startswith('simulation-run-')]
    return max(idx+[-1])
  return -1


def generate_simulation_folder_name(base_path, index=None):
  if index is None:
    index = latest_run_index(base_path)+1
  return f'raw/simulation-run-{int(index):06d}'


class _Named:
  'anything with .Name and .Label (FreeCAD object, source record, scene group)'
  def __init__(self, name, label=None):
    self.Name, self.Label = name, (label if label is not None else name)


def named(x):
  if isinstance(x, dict):
    return _Named(x.get('name', x.get('Name')), x.get('label', x.get('Label')))
  if isinstance(x, (tuple, list)):
    return _Named(*x)
  return x


class SimulationResults:
  def __init__(self, simulationType, basePath, simulationRunFolder=None, flushEverySeconds=5,
               dumpProgressEverySeconds=.2, endAfterIterations=np.inf, endAfterRays=np.inf,
               endAfterHits=np.inf, isMaster=True):
    self.simulationType = simulationType
    self.flushEverySeconds = flushEverySeconds
    self.dumpProgressEverySeconds = dumpProgressEverySeconds
    self._lastFlush = time.time()+self.flushEverySeconds*random.random()
    self._lastDumpedProgress = time.time()+self.dumpProgressEverySeconds*random.random()
    self._lastMasterProgressDump = 0
    self._masterProgressDumpIdx = 0
    self.t0 = time.time()
    self.basePath = str(basePath)
    self.isMaster = isMaster
    if simulationRunFolder is None:
      simulationRunFolder = generate_simulation_folder_name(self.basePath)
    self.simulationRunFolder = simulationRunFolder
    path = f'{self.basePath}/{self.simulationRunFolder}'
    try:
      os.makedirs(path, exist_ok=True)
      if isMaster and not any(f.startswith('uid-') for f in os.listdir(path)):
        with open(f'{path}/uid-{uuid.uuid4()}', 'w'):
          pass
    except Exception:
      raise RuntimeError(f'it seems simulation result path is not writable: {path}')
    self.endAfterIterations = endAfterIterations
    self.endAfterRays = endAfterRays
    self.endAfterHits = endAfterHits
    self.reachedEnd = False
    self.totalIterations = 0
    self.totalTracedRays = 0
    self.totalRecordedRays = 0
    self.totalRecordedHits = 0
    self.hits = None           # list of (source, obj, arrays dict)
    self._bufferedHits = 0
    self.rays = None           # list of (source, [ray dicts])
    self._bufferedRays = 0
    self._cleanedUp = False
    self._lastFingerprintMs = 0
    self.writtenFiles = []
    self._ensureFolderStructureExists()

  # -- folders and names ------------------------------------------------------------------------
  def runFolderPath(self):
    return f'{self.basePath}/{self.simulationRunFolder}'

  def _ensureFolderStructureExists(self):
    for expect in ('raw', 'notebooks'):
      os.makedirs(self.basePath+'/'+expect, exist_ok=True)
    if not os.path.exists(self.basePath+'/README.md'):
      atomic_write_bytes(self.basePath+'/README.md', _README_TEXT.encode())

  def dumpGlobalInfo(self, info):
    atomic_write_bytes(f'{self.runFolderPath()}/global-info.pkl', pickle.dumps(info), overwrite=False)

  def _raiseIfCleanedUp(self):
    if self._cleanedUp:
      raise RuntimeError('this storage was already cleaned up, cannot run requested method')

  def _fingerprint(self, fresh=False):
    'results_store.py:348-350; a fresh one per flush, strictly increasing so two fast flushes never collide'
    if fresh or not self._lastFingerprintMs:
      self._lastFingerprintMs = max(int(time.time()*1e3), self._lastFingerprintMs+1)
    return f'{self._lastFingerprintMs}-pid{os.getpid()}-thread{threading.get_ident()}'

  def _makeFilename(self, kind, source=None, obj=None):
    folder = f'{self.simulationRunFolder}'
    if isinstance(source, str):
      folder += f'/{source}'
    elif source is not None:
      folder += f'/source-{source.Label}'
    if obj is not None:
      folder += f'/object-{obj.Label}'
    os.makedirs(f'{self.basePath}/{folder}', exist_ok=True)
    return f'{self.basePath}/{folder}/{self._fingerprint()}-{kind}.pkl'

  # -- counters -----------------------------------------------------------------------------------
  def incrementRayCount(self, n=1):
    self.totalTracedRays += int(n)

  def incrementIterationCount(self, n=1):
    self.totalIterations += int(n)

  # -- hits ---------------------------------------------------------------------------------------
  # batches of at least this many bytes that the caller only LENDS (views of page-locked engine buffers that the next
  # engine call overwrites) are written to their hit files at once instead of being copied into the buffer first
  DIRECT_WRITE_BYTES = 32 << 20
  ROWS_PER_FILE = 1 << 22            # a lent batch is cut into files of at most this many hits, written by parallel threads
  WRITER_THREADS = 32                # at most; one thread writes ~2 GB/s into a RAM disk (page allocation + copy), so the rate scales with the cores

  def addRayHits(self, source, obj, points, directions, powers, isEntering, metadata=None, borrowed=False):
    '''
    Batch form of addRayHit (results_store.py:641-648): N hits of light source `source` on optical group `obj`.
    metadata: dict key -> array of N values (or (N,3)); only the keys enabled by StoreHit* should be passed.
    borrowed: the arrays are only valid during this call (views of the engine's page-locked delivery buffers): a large
    batch goes straight into hit files (the loader concatenates any number of *-hits.pkl per (source, object),
    results_store.py:74-181), a small one is copied into the buffer.
    '''
    self._raiseIfCleanedUp()
    n = len(powers)
    if n == 0:
      return
    if borrowed:
      if n*57 >= self.DIRECT_WRITE_BYTES:
        self._writeHitsDirect(named(source), named(obj), points, directions, powers, isEntering, metadata or {})
        return
      points, directions, powers = np.array(points, dtype=np.float64), np.array(directions, dtype=np.float64), np.array(powers, dtype=np.float64)
      metadata = {k: np.array(v) for k, v in (metadata or {}).items()}
    entry = dict(points=np.asarray(points, dtype=np.float64).reshape(n, 3),
                 directions=np.asarray(directions, dtype=np.float64).reshape(n, 3),
                 powers=np.asarray(powers, dtype=np.float64).reshape(n),
                 isEntering=np.asarray(isEntering).astype(np.int64).reshape(n))
    for k, v in (metadata or {}).items():
      entry[k] = np.asarray(v)
    if self.hits is None:
      self.hits = []
    self.hits.append((named(source), named(obj), entry))
    self._bufferedHits += n
    self.writeDiskIfNeeded()

  def _writeHitsDirect(self, source, obj, points, directions, powers, isEntering, metadata):
    'hit files of one lent batch, written from the caller\'s arrays without an intermediate copy (pickle protocol 5), in parallel slices'
    from concurrent.futures import ThreadPoolExecutor
    n = len(powers)
    points = np.asarray(points, dtype=np.float64).reshape(n, 3)
    directions = np.asarray(directions, dtype=np.float64).reshape(n, 3)
    powers = np.asarray(powers, dtype=np.float64).reshape(n)
    isEntering = np.asarray(isEntering).reshape(n)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    workers = max(1, min(self.WRITER_THREADS, cores))
    n_files = max(1, -(-n//self.ROWS_PER_FILE))
    if n_files > workers:
      n_files = -(-n_files//workers)*workers          # whole rounds of the thread pool, files of equal size
    rows = max(1, -(-n//n_files))
    slices = [(a, min(n, a+rows)) for a in range(0, n, rows)]
    names = []
    for _ in slices:
      self._fingerprint(fresh=True)
      names.append(self._makeFilename(kind='hits', source=source, obj=obj))

    def write(job):
      (a, b), fname = job
      res = dict(source=source.Name, obj=obj.Name, points=points[a:b], directions=directions[a:b], powers=powers[a:b],
                 isEntering=isEntering[a:b].astype(np.int64))
      for k, v in metadata.items():
        res[k] = np.asarray(v)[a:b]
      with open(fname, 'wb') as f:
        pickle.dump(res, f, protocol=5)           # large array payloads go from the array's memory to the file, no bytes copy
      return fname

    if len(slices) == 1:
      done = [write((slices[0], names[0]))]
    else:
      with ThreadPoolExecutor(max_workers=min(workers, len(slices))) as pool:
        done = list(pool.map(write, zip(slices, names)))
    self.writtenFiles.extend(done)
    self.totalRecordedHits += n

  def addRays(self, source, rays):
    '''
    Batch form of addRay / addSegment / rayComplete (results_store.py:236-257,628-639): `rays` = list of dicts
    dict(points (M+1, 3), powers (M,), media [group Name | None]*M), one per COMPLETE ray of light source `source`.
    '''
    self._raiseIfCleanedUp()
    if not rays:
      return
    if self.rays is None:
      self.rays = []
    self.rays.append((named(source), list(rays)))
    self._bufferedRays += len(rays)
    self.writeDiskIfNeeded()

  def addRayHit(self, source, obj, point, direction, power, isEntering, metadata):
    'single-hit form with the reference signature'
    md = {k: np.asarray([v]) for k, v in (metadata or {}).items()}
    self.addRayHits(source, obj, [point], [direction], [power], [isEntering], md)

  def flush(self):
    'buffered hits -> one pickle per (source, object) with a fresh fingerprint (results_store.py:369-460)'
    self._raiseIfCleanedUp()
    self._fingerprint(fresh=True)
    if self.rays is not None:                                   # results_store.py:380-403: one list of ray dicts per source
      by_file = {}
      for source, rays in self.rays:
        by_file.setdefault(self._makeFilename(kind='rays', source=source), []).extend(rays)
      for fname, dump in by_file.items():
        with open(fname, 'wb') as f:
          pickle.dump(dump, f)
        self.writtenFiles.append(fname)
      self.totalRecordedRays += self._bufferedRays
      self.rays, self._bufferedRays = None, 0
    if self.hits is not None:
      groups = {}
      for source, obj, entry in self.hits:
        fname = self._makeFilename(kind='hits', source=source, obj=obj)
        groups.setdefault(fname, (source, obj, []))[2].append(entry)
      for fname, (source, obj, entries) in groups.items():
        keys = list(entries[0].keys())
        for e in entries[1:]:
          keys += [k for k in e.keys() if k not in keys]
        res = dict(source=source.Name, obj=obj.Name)
        for k in keys:
          parts = []
          for e in entries:
            n = len(e['powers'])
            if k in e:
              parts.append(e[k])
            else:                                   # NaN padding like the reference's metadata handling
              shape = next(x[k].shape[1:] for x in entries if k in x)
              parts.append(np.full((n,)+shape, np.nan))
          res[k] = parts[0] if len(parts) == 1 else np.concatenate(parts, axis=0)
        with open(fname, 'wb') as f:
          pickle.dump(res, f, protocol=5)         # array payloads are written from the arrays' memory (no bytes copy)
        self.writtenFiles.append(fname)
      self.totalRecordedHits += self._bufferedHits
      self.hits, self._bufferedHits = None, 0
    self._lastFlush = time.time() + (.1*random.random()-.05)*self.flushEverySeconds

  # -- progress -----------------------------------------------------------------------------------
  def progressDict(self):
    return dict(simulationType=self.simulationType,
                totalIterations=self.totalIterations,
                totalTracedRays=self.totalTracedRays,
                totalRecordedHits=self.totalRecordedHits+self._bufferedHits,
                totalRecordedRays=self.totalRecordedRays+self._bufferedRays)

  def progressMonitorPath(self):
    return f'{self.runFolderPath()}/progress'

  def dumpProgress(self):
    'results_store.py:462-480: per-worker progress pickle, file name splits on "-" into 4 fields'
    self._raiseIfCleanedUp()
    os.makedirs(self.progressMonitorPath(), exist_ok=True)
    name = f'{self._fingerprint()}-{str(hex(int(random.random()*1e15)))[2:]}.pkl'
    atomic_write_bytes(f'{self.progressMonitorPath()}/{name}', pickle.dumps(self.progressDict()))
    self._lastDumpedProgress = time.time() + (2*random.random()-1)*self.dumpProgressEverySeconds

  def isEndReached(self, progress=None):
    'results_store.py:507-512: strictly greater than the limit'
    p = progress or self.progressDict()
    if (p.get('totalIterations', 0) > self.endAfterIterations
            or p.get('totalTracedRays', 0) > self.endAfterRays
            or p.get('totalRecordedHits', 0) > self.endAfterHits):
      self.reachedEnd = True
    return self.reachedEnd

  def dumpMasterProgress(self, aggregate):
    'results_store.py:515-539: aggregate over all workers (here: all GPU ranks) + the end criteria'
    os.makedirs(self.progressMonitorPath(), exist_ok=True)
    d = dict(simulationType=aggregate.get('simulationType', self.simulationType),
             totalIterations=aggregate.get('totalIterations', 0),
             totalTracedRays=aggregate.get('totalTracedRays', 0),
             totalRecordedHits=aggregate.get('totalRecordedHits', 0),
             totalRecordedRays=aggregate.get('totalRecordedRays', 0),
             endAfterIterations=self.endAfterIterations,
             endAfterRays=self.endAfterRays,
             endAfterHits=self.endAfterHits)
    atomic_write_bytes(f'{self.progressMonitorPath()}/master-{self._masterProgressDumpIdx:09d}', pickle.dumps(d))
    self._masterProgressDumpIdx += 1
    self._lastMasterProgressDump = time.time()
    # like the reference (results_store.py:527-539): master files older than 10 s are removed, the newest always stays
    mine = f'master-{self._masterProgressDumpIdx-1:09d}'
    for name in os.listdir(self.progressMonitorPath()):
      if name.startswith('master-') and name != mine and not name.endswith('.tmp'):
        path = f'{self.progressMonitorPath()}/{name}'
        try:
          if time.time()-os.path.getmtime(path) > 10:
            os.remove(path)
        except OSError:
          pass

  def performanceDescription(self):
    'results_store.py:541-556: the log line benchmark/run-benchmarks.py scrapes'
    dt = max(time.time()-self.t0, 1e-9)
    return f'{self.totalTracedRays/dt:.1e} rays/s, {(self.totalRecordedHits+self._bufferedHits)/dt:.1e} recorded hits/s'

  def writeDiskIfNeeded(self):
    'results_store.py:605-611'
    if time.time()-self._lastFlush > self.flushEverySeconds:
      self.flush()
    if time.time()-self._lastDumpedProgress > self.dumpProgressEverySeconds:
      self.dumpProgress()

  def cleanup(self):
    'remove the progress folder (results_store.py:650-660)'
    import shutil
    shutil.rmtree(self.progressMonitorPath(), ignore_errors=True)
    self._cleanedUp = True
