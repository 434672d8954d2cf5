"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and count per kernel name.
  python tools/launch_summary.py gpurun_out/launches.csv [top]"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors='ignore')) if len(r) > 10]
hdr = rows[0]
iN, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try: v = float(r[iV].replace(',', ''))
    except ValueError: continue
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[iU], 1.0)
    name = re.sub(r'\(.*', '', r[iN])[:110]
    if 'spin_kernel' in name: continue   # torch.cuda._sleep: bench.py's head start for the host in its per-kernel timing phase, not work
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
print(f'{sum(cnt.values())} launches, {T/1e3:.2f} ms total (serialised, cold-cache: compare SHARES; torch.cuda._sleep spins excluded)')
for name, v in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print(f'{v/1e3:9.3f} ms {100*v/T:5.1f}% x{cnt[name]:5d}  {name}')
