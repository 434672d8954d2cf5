"""Developer aid: for each launch of the tcgen05 GEMM kernels in an ncu report (--set full --import-source on), the headline
metrics plus executed counts / stall samples of the barrier waits, TMA loads, MMAs and commits — who waits for whom.
  python tools/ncu_roles.py gpurun_out/x.ncu-rep [launch-index ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
only = [int(a) for a in sys.argv[2:]]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
M = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
     'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
     'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum',
     'launch__shared_mem_per_block_dynamic']
for li, r in enumerate(rows[2:]):
    if only and li not in only: continue
    print(f'==== launch {li}: {r[hdr.index("Kernel Name")][:50]}')
    print('   ' + '  '.join(f'{m.split(".")[0].replace("__", ".")[-22:]}={r[hdr.index(m)]}' for m in M if m in hdr))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', f':::{li + 1}'], capture_output=True, text=True).stdout
    h, seen = None, False
    for q in csv.reader(io.StringIO(src)):
        if q and q[0] == 'Address':
            if seen: break
            seen, h = True, q
            iS, iE, iN = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
            continue
        if h is None or len(q) <= iE: continue
        try: n = int(q[iE] or 0)
        except ValueError: continue
        if n and any(k in q[iS] for k in ('SYNCS.PHASECHK', 'UTCHMMA', 'UTMALDG', 'UTCBAR', 'SYNCS.ARRIVE')):
            print(f'      {q[0][-5:]} {n:9d} {q[iN]:>6s}  {q[iS].strip()[:90]}')
