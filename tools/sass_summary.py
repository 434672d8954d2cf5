"""Per-kernel counts of the Blackwell-specific SASS instructions in libstv.so (cuobjdump -sass): tcgen05 MMA (UTC*MMA), TMEM
loads/stores (LDTM/STTM), TMA (UTMALDG/UTMASTG/UBLKCP), tcgen05 commit/alloc (UTCBAR/UTCATOMSWS), texture gathers (TLD4), packed fp32
(FFMA2/FADD2/FMUL2).   python tools/sass_summary.py [lib] > profiles/r2_sass_summary.txt"""
import collections, re, subprocess, sys
from pathlib import Path
lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent/'slowtv_monodepth_b200'/'libstv.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
WANT = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCOMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'TLD4', 'FFMA2', 'FADD2', 'FMUL2',
        'REDG', 'RED', 'ATOMG', 'MUFU']
fn, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in out.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r'\(.*', '', fn)
        counts[fn] = collections.Counter()
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)', ln)
    if m and fn:
        op, mods = m.group(1), m.group(2)
        total[fn] += 1
        if op in WANT:
            key = op + ('.IM2COL' if 'IM2COL' in mods else '') + ('.2CTA' if '2CTA' in mods else '') + ('.MULTICAST' if 'MULTICAST' in mods else '')
            if op == 'UTMALDG': key = op + ''.join(f'.{d}' for d in re.findall(r'\.(\dD)', mods)) + ('.IM2COL' if 'IM2COL' in mods else '') + ('.2CTA' if '2CTA' in mods else '')
            counts[fn][key] += 1
print(f'# {lib.rsplit("/", 1)[-1]}: {len(counts)} kernels; Blackwell-specific SASS per kernel (cuobjdump -sass, sm_100a)')
for fn, c in counts.items():
    if not c: continue
    print(f'{fn[:100]}  [{total[fn]} instr]')
    print('    ' + '  '.join(f'{k} x{v}' for k, v in sorted(c.items())))
