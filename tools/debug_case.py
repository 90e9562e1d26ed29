import sys, os
sys.path[:0]=['/root/repo','/root/repo/tests']
import numpy as np
import traceray_cases as cases
from freecad.optics_design_workbench_b200 import engine
key=sys.argv[1]
z=np.load('/root/repo/tests/golden/traceray_golden.npz')
g={f.split('/')[1]:z[f] for f in z.files if f.startswith(key+'/')}
build,_=cases.SYNTHETIC_CASES[key]
scene,_,_,settings=build()
cfg=cases.synthetic_cfg(settings, record_all_hits=True, wavelength=500.0, hit_capacity=len(g['hit_powers'])+1000)
eng=engine.Engine(0)
ds=eng.scene(scene)
with ds.trace_rays(cfg,g['origins'],g['directions']) as res:
  hits,summary=res.hits(sort=True),res.ray_summary()
print('n hits gpu',len(hits['powers']),'golden',len(g['hit_powers']))
ns=np.diff(g['seg_offsets'])
bad=np.nonzero(summary['n_segments']!=ns)[0]
print('rays with different segment counts',bad[:20], len(bad))
# per-ray compare
for r in range(len(ns)):
  a=hits['face_id'][hits['ray_index']==r]; b=g['hit_face_id'][g['hit_ray']==r]
  if len(a)!=len(b) or (a!=b).any():
    print('ray',r,'gpu',a,'ref',b)
    pa=hits['points'][hits['ray_index']==r]; pb=g['hit_points'][g['hit_ray']==r]
    print(' gpu pts',pa[:6]); print(' ref pts',pb[:6])
    print(' gpu grp',hits['group'][hits['ray_index']==r],' ref grp',g['hit_group'][g['hit_ray']==r])
    print(' faces kinds', scene.faces['kind'], scene.faces['face_id'], scene.faces['group'])
    break
