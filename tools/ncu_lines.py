"""Per-CUDA-source-line instruction and stall-sample totals of one kernel from an ncu report captured with --import-source on.
  python tools/ncu_lines.py gpurun_out/x.ncu-rep kernel-regex [top]"""
import collections, csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--kernel-name', f'regex:{rx}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
inst, samp, text = collections.Counter(), collections.Counter(), {}
h, f, seen = None, '', set()
for r in rows:
    if r and r[0] == 'File Path':
        f = r[1].rsplit('/', 1)[-1]
        continue
    if r and r[0] == 'Function Name':
        key = (f, r[1])
        if key in seen: break  # first launch of the kernel only
        seen.add(key)
        continue
    if r and r[0] == 'Line No':
        h = r; iE, iN = h.index('Instructions Executed'), h.index('# Samples')
        continue
    if h is None or len(r) <= iE or not r[0].strip().isdigit(): continue
    k = (f, int(r[0]))
    text[k] = r[1].strip()
    try: inst[k] += int(r[iE] or 0); samp[k] += int(r[iN] or 0)
    except ValueError: pass
ti, ts = sum(inst.values()), sum(samp.values())
print(f'{ti} warp instructions, {ts} samples')
for k, n in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
    print(f'{k[0][:16]:16s}:{k[1]:4d} {100*inst[k]/max(ti,1):5.1f}% inst {100*n/max(ts,1):5.1f}% samples  {text.get(k, "")[:100]}')
