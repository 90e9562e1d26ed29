'''Developer script: quick CUDA-vs-oracle parity and timing on every scene fixture (run under gpurun).'''
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import numpy as np
from freecad.optics_design_workbench_b200 import engine
from freecad.optics_design_workbench_b200.simulation.setup import prepare
from oracle import Oracle

def main():
  eng = engine.Engine(0)
  print(eng.device_name())
  orc = Oracle()
  seed = 0x0DDB1A5E
  nbig = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
  for name in ['minimal', 'lensesAndMirrors', 'lensesAndMirrorsSequential', 'hugeArray']:
    sim = prepare(os.path.join(ROOT, 'tests', 'golden', 'scenes', name+'.npz'))
    ds = eng.scene(sim.scene)
    sa = sim.source_args(0)
    dsrc = eng.source(sa)
    # sampler parity
    n = 50000
    g = dsrc.sample(seed, 123, n)
    o = orc.sample_mc(sa, seed, 123, n)
    print(f'[{name}] sampler: max|dtheta|={np.abs(g["first"]-o["first"]).max():.2e} max|dphi|={np.abs(g["phi"]-o["phi"]).max():.2e} '
          f'max|ddir|={np.abs(g["directions"]-o["directions"]).max():.2e}')
    # MC parity with every intersection recorded
    n = 20000
    cfg = sim.cfg(record_all_hits=True, hit_capacity=n*int(sim.settings['MaxIntersections']) if name == 'hugeArray' else n*10)
    t0 = time.time()
    ref = orc.trace_mc(sim.scene, sa, cfg, seed, 0, n, hit_capacity=int(cfg.cfg.hit_capacity), threads=0)
    t_or = time.time()-t0
    with ds.trace_mc(dsrc, cfg, seed, 0, n) as res:
      c = res.counts; h = res.hits(sort=True)
    rh = ref['hits']
    same_len = len(h['group']) == len(rh['group'])
    seq_ok = same_len and np.array_equal(h['group'], rh['group']) and np.array_equal(h['ray_index'], rh['ray_index']) and np.array_equal(h['face_id'], rh['face_id'])
    print(f'[{name}] MC {n} rays: gpu {c} ')
    print(f'[{name}]            oracle {ref["counts"]} ({t_or:.2f}s)')
    if seq_ok:
      print(f'[{name}] sequences identical; max|dP|={np.abs(h["points"]-rh["points"]).max():.3e} max|dD|={np.abs(h["directions"]-rh["directions"]).max():.3e} '
            f'entering equal={np.array_equal(h["is_entering"], rh["is_entering"])}')
    else:
      # per-ray comparison
      bad = 0
      import collections
      def per_ray(hh):
        d = collections.defaultdict(list)
        for r, f in zip(hh['ray_index'], hh['face_id']): d[int(r)].append(int(f))
        return d
      a, b = per_ray(h), per_ray(rh)
      diff = [r for r in range(n) if a.get(r) != b.get(r)]
      print(f'[{name}] SEQUENCE MISMATCH on {len(diff)} of {n} rays; first: {diff[:5]}')
      for r in diff[:3]:
        print('   gpu   ', a.get(r)); print('   oracle', b.get(r))
    # timing
    cfg = sim.cfg(hit_capacity=2*nbig)
    for rep in range(3):
      with ds.trace_mc(dsrc, cfg, seed, 0, nbig) as res:
        c = res.counts; ms = res.kernel_ms
    print(f'[{name}] {nbig} rays: {ms:.2f} ms  -> {c["segments"]/ms*1e3:.3e} segments/s, {nbig/ms*1e3:.3e} rays/s, seg/ray={c["segments"]/nbig:.2f}, hits={c["hits"]}, dropped={c["hits_dropped"]}')
    cfg = sim.cfg(store_hits=False)
    with ds.trace_mc(dsrc, cfg, seed, 0, nbig) as res:
      c = res.counts; ms = res.kernel_ms
    print(f'[{name}] no hit store: {ms:.2f} ms -> {c["segments"]/ms*1e3:.3e} segments/s')

if __name__ == '__main__':
  main()
