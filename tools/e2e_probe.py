'''Developer script: what bounds the end-to-end (hit lists to host) path?  raw pinned D2H bandwidth vs odw_trace_mc_host.'''
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import numpy as np, torch
from freecad.optics_design_workbench_b200 import engine
from freecad.optics_design_workbench_b200.simulation.setup import prepare

def main():
  n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
  d = torch.empty(1 << 30, dtype=torch.uint8, device='cuda')
  h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
  for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d, non_blocking=True); torch.cuda.synchronize()
    print(f'raw pinned D2H 1 GiB: {(1 << 30)/(time.perf_counter()-t0)/1e9:.1f} GB/s', flush=True)
  eng = engine.Engine(0)
  sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', 'lensesAndMirrors.npz'))
  ds, dsrc = eng.scene(sim.scene), eng.source(sim.source_args(0))
  cap = int(n*1.05)+1024
  for columns in (('points', 'directions', 'powers', 'is_entering'), ('points',), ('powers',)):
    arrays, view = eng.pinned_hit_arrays(cap, columns)
    per_hit = sum(dict(points=24, directions=24, powers=8, is_entering=1)[c] for c in columns)
    for chunk in (1 << 21, 1 << 23, 1 << 25):
      os.environ['ODW_HOST_CHUNK'] = str(chunk)
      best = 1e9
      for rep in range(3):
        t0 = time.perf_counter()
        c, got = ds.trace_mc_host(dsrc, sim.cfg(store_hits=True), 1, rep*n, n, view)
        best = min(best, time.perf_counter()-t0)
      print(f'columns={columns} chunk=2^{chunk.bit_length()-1}: {best*1e3:.1f} ms -> {c["segments"]/best:.3e} seg/s, D2H {got*per_hit/best/1e9:.1f} GB/s', flush=True)

if __name__ == '__main__':
  main()
