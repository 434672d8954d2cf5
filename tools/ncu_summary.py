"""Summarise an ncu report: key metrics per kernel + SASS opcode mix + stall-sample distribution.
  python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tex.sum', 'l1tex__texin_requests.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print('----')
    for i in idx: print(f'  {hdr[i]} = {r[i]} {rows[1][i]}')
args = ['ncu', '-i', rep, '--page', 'source', '--csv']
if rx: args += ['--kernel-name', f'regex:{rx}']
src = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
ops, samp = collections.Counter(), collections.Counter()
stall_cols = {}
tot = tots = 0
per_kernel = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        if tot: per_kernel.append((name, tot, tots, ops, samp, stall))
        name = r[1]; ops, samp, stall = collections.Counter(), collections.Counter(), collections.Counter(); tot = tots = 0; h = None
        continue
    if r and r[0] == 'Address':
        h = r; iS, iE, iN = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
        stall_cols = {i: c for i, c in enumerate(h) if c.startswith('stall_')}
        continue
    if h is None or len(r) <= iE: continue
    toks = r[iS].strip().split()
    if not toks: continue
    op = (toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]).split('.')[0]
    n = int(r[iE] or 0); ops[op] += n; tot += n
    s = int(r[iN] or 0); samp[op] += s; tots += s
    for i, c in stall_cols.items():
        try: stall[c] += int(r[i] or 0)
        except ValueError: pass
if tot: per_kernel.append((name, tot, tots, ops, samp, stall))
for name, tot, tots, ops, samp, stall in per_kernel:
    print(f'==== {name[:80]}: {tot} warp instr, {tots} samples')
    for op, n in ops.most_common(18): print(f'   {op:10s} {n:12d} {100*n/tot:5.1f}%  samples {100*samp[op]/max(tots,1):5.1f}%')
    st = sum(stall.values())
    print('   stalls:', ', '.join(f'{k[6:]} {100*v/max(st,1):.0f}%' for k, v in stall.most_common(8)))
