'''Developer script: join an ncu report's per-SASS-instruction counters with nvdisasm line info.
usage: ncu_lines.py report.ncu-rep lib.so kernel_mangled_substring [srcdir] [top]
Prints the instruction mix, the stall reasons and the hottest source lines of the first captured launch.'''
import collections, csv, os, re, subprocess, sys, tempfile

def main():
  rep, lib, kern = sys.argv[1:4]
  srcdir = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'freecad', 'optics_design_workbench_b200', 'csrc')
  top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
  tmp = tempfile.mkdtemp()
  subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
  cubin = [f for f in os.listdir(tmp) if f.startswith(os.environ.get('ODW_CUBIN', 'odw_kernels')+'.')][0]
  sass = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
  start = [i for i, l in enumerate(sass) if l.startswith('.text.') and kern in l][0]
  end = next((i for i in range(start+1, len(sass)) if sass[i].startswith('//--------------------- .text')), len(sass))
  seq, cur = [], None
  for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
      cur = (os.path.basename(m.group(1)), int(m.group(2)))
      continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
      seq.append((m.group(2), cur))
  out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
  rows, hdr, b, nk = list(csv.reader(out.split('\n'))), None, [], 0
  for r in rows:
    if r and r[0] == 'Kernel Name':
      nk += 1
      if nk > 1:
        break
      continue
    if r and r[0] == 'Address':
      hdr = r
      continue
    if hdr and len(r) > 5:
      b.append(r)
  assert len(b) == len(seq), (len(b), len(seq), 'the .so does not match the profiled build')
  iE, iS = hdr.index('Instructions Executed'), hdr.index('# Samples')
  tot, tots = sum(int(r[iE]) for r in b), sum(int(r[iS]) for r in b)
  print(f'static instructions {len(b)}, executed warp instructions {tot}, samples {tots}')
  op, ops = collections.Counter(), collections.Counter()
  for r in b:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1])
    o = m.group(2).split('.')[0] if m else '?'
    op[o] += int(r[iE]); ops[o] += int(r[iS])
  for o, c in op.most_common(24):
    print(f'  {o:10s} {c/tot*100:6.2f}% executed {ops[o]/tots*100:6.2f}% samples')
  for name in hdr:
    if name.startswith('stall_') and 'Not Issued' not in name:
      s = sum(int(r[hdr.index(name)]) for r in b)
      if s/tots > 0.01:
        print(f'  {name:28s} {s/tots*100:5.1f}%')
  only = os.environ.get('ODW_OPCODES')      # e.g. ODW_OPCODES=LDL,STL: attribute only these opcodes to source lines
  if only:
    keep = set(only.split(','))
    pairs = [(sc, r) for sc, r in zip(seq, b) if (re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[1]) or [None]*3)[2] in keep]
    seq, b = [x[0] for x in pairs], [x[1] for x in pairs]
  iT = hdr.index('Thread Instructions Executed')
  byline, bys, byt = collections.Counter(), collections.Counter(), collections.Counter()
  for (ins, c), r in zip(seq, b):
    byline[c] += int(r[iE]); bys[c] += int(r[iS]); byt[c] += int(r[iT])
  src = {}
  # per function: a line belongs to the last `__device__` / `__global__` definition that starts at or before it
  byfun, byfuns, byfunt = collections.Counter(), collections.Counter(), collections.Counter()
  starts = {}
  for k, c in byline.items():
    f, l = k if k else ('?', 0)
    if f not in starts:
      try:
        lines = open(os.path.join(srcdir, f)).read().split('\n')
      except OSError:
        lines = []
      starts[f] = [(i+1, re.sub(r'\(.*', '', ln.split('(')[0]).split()[-1]) for i, ln in enumerate(lines)
                   if re.match(r'\s*(template\s*<[^>]*>\s*)?(static\s+)?__(device|global)__', ln) and '(' in ln]
    name = '?'
    for l0, n in starts[f]:
      if l0 <= l:
        name = n
    byfun[(f, name)] += c; byfuns[(f, name)] += bys[k]; byfunt[(f, name)] += byt[k]
  print('by function (inlined code is attributed to the function it was written in):')
  for (f, name), c in byfun.most_common(25):
    print(f'  {f[:16]:16s} {name[:28]:28s} {c/tot*100:5.2f}% instr {byfuns[(f, name)]/tots*100:5.2f}% samples {byfunt[(f, name)]/max(c, 1):5.1f} lanes')
  for k, c in byline.most_common(top):
    f, l = k if k else ('?', 0)
    if f not in src:
      try:
        src[f] = open(os.path.join(srcdir, f)).read().split('\n')
      except OSError:
        src[f] = []
    t = src[f][l-1].strip()[:100] if 0 < l <= len(src[f]) else ''
    print(f'{f[:15]:15s} {l:4d} {c/tot*100:5.2f}% instr {bys[k]/tots*100:5.2f}% samples {byt[k]/max(c, 1):5.1f} lanes | {t}')

if __name__ == '__main__':
  main()
