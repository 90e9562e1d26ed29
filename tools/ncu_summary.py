'''Developer script: condense an `ncu --set full` report into the small files committed under profiles/.
usage: ncu_summary.py report.ncu-rep out_prefix [--traffic]
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
        'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg.per_second']

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
    d = out[0]
    rd = d['dram__bytes_read.sum']['value']*scale[d['dram__bytes_read.sum']['unit']]
    wr = d['dram__bytes_write.sum']['value']*scale[d['dram__bytes_write.sum']['unit']]
    path = os.path.join(os.path.dirname(os.path.abspath(prefix)), 'traffic.json')
    with open(path, 'w') as f:
      json.dump(dict(kernel=d['kernel'], dram_bytes_per_launch=rd+wr, dram_read_bytes=rd, dram_write_bytes=wr,
                     duration_ms=d['gpu__time_duration.sum']['value']*(1e-3 if d['gpu__time_duration.sum']['unit'] == 'us' else 1 if d['gpu__time_duration.sum']['unit'] == 'ms' else 1e-6),
                     source=os.path.basename(rep), rays_per_launch=2097152,
                     launch='2^21 Monte-Carlo rays of lensesAndMirrors, hit lists stored (ODW_RAYS_PER_LAUNCH=2097152: the default 2^18-ray '
                            'launch leaves most of its hit list dirty in the 126 MB L2 when the kernel ends, which hides the writes from the counter)'), f, indent=1)
  print('wrote', prefix+'_details.csv', prefix+'_raw.json')

if __name__ == '__main__':
  main()
