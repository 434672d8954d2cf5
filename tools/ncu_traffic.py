"""DRAM traffic of the photometric kernels from an `ncu --set full` report -> profiles/r2_photo_traffic.json (read by bench.py's
`roofline.traffic`).  python tools/ncu_traffic.py gpurun_out/x.ncu-rep [kernel-regex] [label]"""
import csv, io, json, re, subprocess, sys
from pathlib import Path
rep = sys.argv[1]
rx = re.compile(sys.argv[2] if len(sys.argv) > 2 else 'photo_fused|photo_error|pull_|fused_finalize|fused_loss_reduce')
label = sys.argv[3] if len(sys.argv) > 3 else Path(rep).name
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {n: hdr.index(n) for n in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum')}
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}
per = {}
for r in rows[2:]:
    name = r[col['Kernel Name']]
    if not rx.search(name): continue
    short = re.sub(r'<.*', '', name.split('(')[0]).replace('stv::', '')
    d = per.setdefault(short, {'launches': 0, 'read': 0.0, 'write': 0.0, 'us': 0.0})
    d['launches'] += 1
    for k, c in (('read', 'dram__bytes_read.sum'), ('write', 'dram__bytes_write.sum'), ('us', 'gpu__time_duration.sum')):
        d[k] += float(r[col[c]].replace(',', ''))*scale.get(units[col[c]], 1)
out = {'source': f'ncu --set full, {label}: dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the captured launches',
       'kernels': {k: {'launches': v['launches'], 'dram_read_bytes': v['read']/v['launches'], 'dram_write_bytes': v['write']/v['launches'],
                       'gpu_time_us_under_ncu': v['us']/v['launches']} for k, v in per.items()}}
out['dram_bytes_per_step'] = sum(v['dram_read_bytes'] + v['dram_write_bytes'] for v in out['kernels'].values())
Path('profiles').mkdir(exist_ok=True)
Path('profiles/r2_photo_traffic.json').write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
