'''Developer script: kernel time vs ray count / hit storage for one scene (run under gpurun).'''
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from freecad.optics_design_workbench_b200 import engine
from freecad.optics_design_workbench_b200.simulation.setup import prepare

def main():
  name = sys.argv[1] if len(sys.argv) > 1 else 'lensesAndMirrors'
  sizes = [float(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [4e6, 2e7, 1e8]
  eng = engine.Engine(0)
  sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', name+'.npz'))
  ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
  for n in sizes:
    n = int(n)
    for store in (True, False):
      cfg = sim.cfg(store_hits=store, hit_capacity=int(1.05*n)+1024)
      best = 1e9
      for rep in range(4):
        with ds.trace_mc(dsrc, cfg, 0x0DDB1A5E, rep*n, n) as res:
          c, ms, mhz = res.counts, res.kernel_ms, res.sm_clock_mhz
        best = min(best, ms)
      print(f'{name} n={n:.1e} store={store}: best {best:.2f} ms, last {ms:.2f} ms -> {c["segments"]/best*1e3:.3e} seg/s ({best*1e6/n:.2f} ns/ray) sm {mhz:.0f} MHz', flush=True)

if __name__ == '__main__':
  main()
