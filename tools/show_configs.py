'''Developer script: one line per bench.py JSON line of the given .jsonl files.'''
import json, sys
for f in sys.argv[1:]:
  for line in open(f):
    line = line.strip()
    if not line.startswith('{'):
      continue
    d = json.loads(line)
    e, p = d.get('e2e', {}), d.get('e2e_plugin', {})
    ar = d.get('allreduce') or {}
    print(f"N={d['n_gpus']} {d['config']['workload'][:60]:60s} value {d['value']:.3e} ms/step {d['ms_per_step']:.2f} e2e {e.get('value', 0):.3e} "
          f"plugin {p.get('value', 0):.2e} frac {d['roofline']['frac']:.3f} allreduce_ms {ar.get('ms_per_step')} cpu {d.get('cpu_baseline', {}).get('value')}")
