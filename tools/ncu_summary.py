'''Developer script: condense an `ncu --set full` report into the small files committed under profiles/.
usage: ncu_summary.py report.ncu-rep out_prefix [--traffic [scene rays_per_launch segments_per_launch]]
Writes <out_prefix>_details.csv (ncu --page details) and <out_prefix>_raw.json (selected raw counters per captured
launch); with --traffic also profiles/traffic.json (dram bytes per launch of the first capture, read by bench.py).'''
import csv, io, json, os, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg.per_second',
        'sm__cycles_elapsed.avg', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed',
        'sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'smsp__warps_eligible.avg.per_cycle_active']

def main():
  rep, prefix = sys.argv[1], sys.argv[2]
  det = subprocess.run(['ncu', '-i', rep, '--page', 'details', '--csv'], capture_output=True, text=True).stdout
  with open(prefix+'_details.csv', 'w') as f:
    f.write(det)
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units = rows[0], rows[1]
  out = []
  for r in rows[2:]:
    d = {'kernel': r[hdr.index('Kernel Name')]}
    for w in WANT:
      if w in hdr:
        i = hdr.index(w)
        try:
          d[w] = dict(value=float(r[i].replace(',', '')), unit=units[i])
        except ValueError:
          d[w] = dict(value=r[i], unit=units[i])
    out.append(d)
  with open(prefix+'_raw.json', 'w') as f:
    json.dump(out, f, indent=1)
  if '--traffic' in sys.argv and out:
    scale = dict(byte=1, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9)
    extra = sys.argv[sys.argv.index('--traffic')+1:]
    scene = extra[0] if extra else 'lensesAndMirrors'
    rays = int(extra[1]) if len(extra) > 1 else 2097152
    segments = float(extra[2]) if len(extra) > 2 else None
    d = out[0]
    val = lambda k: d[k]['value'] if k in d else None
    rd = d['dram__bytes_read.sum']['value']*scale[d['dram__bytes_read.sum']['unit']]
    wr = d['dram__bytes_write.sum']['value']*scale[d['dram__bytes_write.sum']['unit']]
    t = d['gpu__time_duration.sum']
    rec = dict(kernel=d['kernel'], dram_bytes_per_launch=rd+wr, dram_read_bytes=rd, dram_write_bytes=wr,
               duration_ms=t['value']*(1e-3 if t['unit'] == 'us' else 1 if t['unit'] == 'ms' else 1e-6 if t['unit'] == 'ns' else 1e3),
               source=os.path.basename(rep), rays_per_launch=rays, segments_per_launch=segments,
               launch=f'{rays} Monte-Carlo rays of {scene} in one launch, hit lists stored')
    cyc = val('sm__cycles_elapsed.avg')
    ops = [val('smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % o) for o in ('dfma', 'dmul', 'dadd')]
    if cyc and all(o is not None for o in ops):
      rec['fp64_flop_per_launch'] = (2*ops[0] + ops[1] + ops[2])*cyc      # thread-level DFMA x 2 + DMUL + DADD
      rec['fp64_thread_inst_per_cycle'] = dict(dfma=ops[0], dmul=ops[1], dadd=ops[2], peak_dfma=val('sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained'))
    rec['fp64_pipe_active_pct'] = val('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')
    rec['issue_active_pct'] = val('smsp__issue_active.avg.pct_of_peak_sustained_active')
    name = 'traffic.json' if scene == 'lensesAndMirrors' else f'traffic_{scene}.json'
    path = os.path.join(os.path.dirname(os.path.abspath(prefix)), name)
    with open(path, 'w') as f:
      json.dump(rec, f, indent=1)
  print('wrote', prefix+'_details.csv', prefix+'_raw.json')

if __name__ == '__main__':
  main()
