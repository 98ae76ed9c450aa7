"""Per CUDA source line: instructions executed and stall samples of every kernel in an .ncu-rep.
    python profiles/ncu_lines.py <rep> [top-N]"""
import csv, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
fn, hdr, rows = None, None, {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == 'Function Name':
        fn = r[1].split('(')[0].split('::')[-1]
        rows.setdefault(fn, [])
    elif r[0] == 'Line No':
        hdr = r
    elif fn and hdr and r[0].isdigit() and len(r) == len(hdr):
        rows[fn].append(r)
for fn, rs in rows.items():
    isamp, ie, it = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    tot_i = sum(int(r[ie]) for r in rs) or 1
    tot_s = sum(int(r[isamp]) for r in rs) or 1
    print('=' * 10, fn, 'warp-instr', tot_i, 'samples', tot_s)
    for r in sorted(rs, key=lambda r: -int(r[isamp]))[:topn]:
        e = int(r[ie]) or 1
        print(f"{int(r[isamp]) / tot_s * 100:5.1f}%smp {int(r[ie]) / tot_i * 100:5.1f}%ins thr={int(r[it]) / e:4.1f} L{r[0]:>5s} {r[1].strip()[:110]}")
