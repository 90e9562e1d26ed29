'''
Orchestration swap: tools/odw-gpu-worker stands in for the `FreeCAD -c` children of the reference's
worker pool (reference simulation/processes/worker_process.py:65-69,139-184; simulation_loop.py:450-507).

  * protocol, here (CPU, needs /root/reference): the reference's OWN WorkerProcess class starts the GPU worker through
    $APPIMAGE, probes it (isBusy: echo of a random number) and sends it its simulation text; without a GPU the worker reports
    the missing device on stderr and keeps serving the console, like a FreeCAD child whose statement raised.
  * a real run, on the GPU box: the text of worker_process.py:151-163 is written to the worker, a stand-in master polls the
    progress files the way results_store.getProgressByWorker does (:558-603), drops `simulation-is-done`, and the hit files
    in the run folder hold exactly the hits the worker reported.
'''
import glob
import os
import pickle
import shutil
import subprocess
import sys
import time
import types

import numpy as np
import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, 'tools', 'odw-gpu-worker')


def master_text(fcstd_path, action, run_folder, parent_pid, jupyter=True):
  'what WorkerProcess.startSimulation writes to the child (worker_process.py:151-163), character for character'
  return (f'\r\n'
          f'for doc in App.listDocuments():\r\n'
          f'  App.closeDocument(doc)'+'\r\n'*3+
          f'App.openDocument({repr(fcstd_path)})\r\n'
          f'from freecad.optics_design_workbench import simulation\r\n'
          f'simulation.setIsJupyterContext({repr(jupyter)})\r\n'
          f'simulation.runSimulation('
          f'action={repr(action)}, '
          f'slaveInfo=dict(simulationRunFolder={repr(run_folder)}, '
          f'               parentPid={parent_pid}))\r\n'
          +f'\r\n'*3+
          f'for doc in App.listDocuments():\r\n'
          f'  App.closeDocument(doc)'+'\r\n'*3)


def probe(proc, timeout=60.0):
  'WorkerProcess.isBusy (worker_process.py:165-184): ask the child to print a random number; True = it answered'
  token = f'{np.random.random()}'
  proc.stdin.write(f'\r\nprint("{token}")\r\n\r\n'); proc.stdin.flush()
  t0 = time.time()
  while time.time()-t0 < timeout:
    line = proc.stdout.readline()
    if token in line:
      return True
  return False


def test_console_protocol_without_a_simulation():
  'echo probe, document bookkeeping and the import the master\'s text does — no device needed'
  p = subprocess.Popen([WORKER, '-c'], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
  try:
    assert probe(p)
    p.stdin.write('\r\nApp.openDocument("/tmp/some-project.FCStd")\r\nfrom freecad.optics_design_workbench import simulation\r\n'
                  'simulation.setIsJupyterContext(True)\r\nprint(sorted(App.listDocuments()))\r\n'
                  'for doc in App.listDocuments():\r\n  App.closeDocument(doc)\r\n\r\n\r\nprint(len(App.listDocuments()))\r\n')
    p.stdin.flush()
    seen = []
    while len(seen) < 20 and '0' not in seen:             # an interactive console also echoes expression values (so does FreeCAD's)
      seen.append(p.stdout.readline().strip())
    assert "['some-project']" in seen and seen[-1] == '0'
    assert probe(p)
  finally:
    p.stdin.close(); p.wait(timeout=20)


@pytest.mark.reference
def test_reference_worker_process_drives_the_gpu_worker(tmp_path, monkeypatch):
  import reference_shim
  reference_shim.load()
  procs = sys.modules['odw_ref.simulation.processes']
  procs.__path__ = [reference_shim.ROOT+'/simulation/processes']
  fe = types.ModuleType('odw_ref.simulation.freecad_elements'); fe.keepGuiResponsive = lambda: None
  sys.modules['odw_ref.simulation.freecad_elements'] = fe; sys.modules['odw_ref.simulation'].freecad_elements = fe
  fe2 = types.ModuleType('odw_ref.freecad_elements'); fe2.keepGuiResponsive = lambda: None
  sys.modules.setdefault('odw_ref.freecad_elements', fe2)
  if not hasattr(sys.modules['odw_ref.freecad_elements'], 'keepGuiResponsive'):
    sys.modules['odw_ref.freecad_elements'].keepGuiResponsive = lambda: None
  monkeypatch.setenv('APPIMAGE', WORKER)
  from odw_ref.simulation.processes import worker_process
  doc = tmp_path/'minimal.FCStd'
  shutil.copyfile('/root/reference/benchmark/minimal.FCStd', doc)
  procs.simulatingDocument = lambda: types.SimpleNamespace(getFileName=lambda: str(doc))
  w = worker_process.WorkerProcess(isJupyterContext=True)
  try:
    assert w.isRunning()
    for _ in range(30):                                   # the interpreter needs a moment to come up; then the probe is answered
      if not w.isBusy():
        break
    else:
      raise AssertionError('the worker never answered the liveness probe')
    w.startSimulation('true', 'raw/simulation-run-000000')
    deadline = time.time()+60
    while w.isBusy() and time.time() < deadline:          # busy while runSimulation runs (on a GPU box: until the flag is dropped)
      open(tmp_path/'minimal.OpticsDesign'/'simulation-is-done', 'w').close() if (tmp_path/'minimal.OpticsDesign').exists() else None
    assert not w.isBusy() and w.isRunning()               # it came back to the console (without a GPU: after reporting the missing device)
  finally:
    w.terminate()


@pytest.mark.gpu
def test_gpu_worker_runs_a_simulation_for_a_master(tmp_path):
  src = os.path.join(ROOT, 'baseline', '_ref', 'scenes', 'lensesAndMirrors.FCStd')
  if not os.path.exists(src):
    pytest.skip('the reference\'s benchmark documents did not travel (baseline/_ref/scenes)')
  doc = str(tmp_path/'lensesAndMirrors.FCStd')
  shutil.copyfile(src, doc)
  base = str(tmp_path/'lensesAndMirrors.OpticsDesign')
  run = 'raw/simulation-run-000000'
  os.makedirs(f'{base}/{run}/progress')
  open(f'{base}/simulation-is-running', 'w').close()
  p = subprocess.Popen([WORKER, '-c'], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
  try:
    assert probe(p)
    p.stdin.write(master_text(doc, 'true', run, os.getpid())); p.stdin.flush()
    latest, t0 = {}, time.time()
    while time.time()-t0 < 120 and latest.get('totalTracedRays', 0) < 3e6:     # the master's view: latest progress file of the worker
      for f in sorted(glob.glob(f'{base}/{run}/progress/*.pkl')):
        try:
          latest = pickle.load(open(f, 'rb'))
        except Exception:
          pass
      time.sleep(0.05)
    assert latest.get('totalTracedRays', 0) >= 3e6, (latest, p.stderr.read() if p.poll() is not None else '')
    open(f'{base}/simulation-is-done', 'w').close()                          # end criterion reached: the master ends the run
    assert probe(p, timeout=120)                                             # the worker left runSimulation and answers again
    for f in sorted(glob.glob(f'{base}/{run}/progress/*.pkl')):
      latest = pickle.load(open(f, 'rb'))
    files = glob.glob(f'{base}/{run}/source-*/object-*/*-hits.pkl')
    assert files
    n_hits = sum(len(pickle.load(open(f, 'rb'))['powers']) for f in files)
    assert n_hits == latest['totalRecordedHits'] and latest['totalTracedRays'] >= n_hits > 0.9*latest['totalTracedRays']
    one = pickle.load(open(files[0], 'rb'))
    assert set(one) >= {'source', 'obj', 'points', 'directions', 'powers', 'isEntering'} and one['points'].shape[1] == 3
  finally:
    try:
      p.stdin.close()
    except Exception:
      pass
    p.wait(timeout=60)
