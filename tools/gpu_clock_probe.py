'''Developer script: does the trace kernel slow down under sustained load (clock / power management)?'''
import os, sys, subprocess, threading, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from freecad.optics_design_workbench_b200 import engine
from freecad.optics_design_workbench_b200.simulation.setup import prepare

def main():
  eng = engine.Engine(0)
  sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'lensesAndMirrors.npz'))
  ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
  rows = []
  proc = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active', '--format=csv,noheader,nounits', '-lms', '20'],
                          stdout=subprocess.PIPE, text=True)
  def rd():
    for line in proc.stdout: rows.append((time.perf_counter(), line.strip()))
  threading.Thread(target=rd, daemon=True).start()
  n = 4_000_000
  cfg = sim.cfg(store_hits=False)
  times = []
  t0 = time.perf_counter()
  for rep in range(80):
    with ds.trace_mc(dsrc, cfg, 1, rep*n, n) as res:
      times.append((time.perf_counter()-t0, res.kernel_ms))
  print('per-launch ms:', ' '.join(f'{ms:.2f}' for _, ms in times))
  n = 100_000_000
  t1 = time.perf_counter()
  with ds.trace_mc(dsrc, cfg, 1, 0, n) as res:
    print('1e8:', res.kernel_ms, 'ms')
  t2 = time.perf_counter()
  time.sleep(0.3)
  proc.terminate()
  print('clock samples (t, sm MHz, mem MHz, W, C, reasons):')
  for t, r in rows:
    print(f'  {t-t0:7.3f} {r}' + ('   <- during 1e8 launch' if t1 <= t <= t2 else ''))

if __name__ == '__main__':
  main()
