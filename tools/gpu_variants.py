'''Developer script: the same workload through several builds of the library (ODW_LIB), one subprocess each.
usage: gpu_variants.py scene n_rays lib1.so lib2.so ...   (run under gpurun; prints best-of-4 segments/s per build)'''
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
scene, n = sys.argv[1], sys.argv[2]
for rep in range(2):
  for lib in sys.argv[3:]:
    env = dict(os.environ, ODW_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'gpu_sweep.py'), scene, n], env=env, capture_output=True, text=True)
    for line in out.stdout.splitlines():
      if 'store=True' in line:
        print(os.path.basename(lib), line, flush=True)
    if out.returncode:
      print(os.path.basename(lib), 'FAILED', out.stderr[-400:], flush=True)
