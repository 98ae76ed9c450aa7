"""Summarise an .ncu-rep here (no GPU needed): python profiles/summarize_ncu.py <rep> [top-N source lines]
Prints, per profiled launch, the metrics the roofline discussion uses and the hottest SASS lines."""
import csv, subprocess, sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__average_warps_issue_stalled',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared', 'lts__t_bytes.sum ', 'l1tex__m_xbar2l1tex_read_bytes.sum ', 'sm__throughput.avg.pct',
        'gpu__dram_throughput.avg.pct', 'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct', 'smsp__inst_executed_op_shared_atom']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
kn = hdr.index('Kernel Name')
for li, vals in enumerate(rows[2:]):
    print('=' * 20, 'launch', li, vals[kn][:80])
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k) or h == k.strip() for k in KEYS):
            if 'stalled' in h and float(v or 0) < 0.15:
                continue
            print(f"{h:85s} {u:12s} {v}")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-kernel-base', 'function'], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if 'Source' in r and 'Instructions Executed' in r:
        cur = {'hdr': r, 'rows': []}
        blocks.append(cur)
    elif cur is not None and len(r) == len(cur['hdr']):
        cur['rows'].append(r)
for bi, b in enumerate(blocks):
    h = b['hdr']
    ia, ie, it, isamp = h.index('Source'), h.index('Instructions Executed'), h.index('Avg. Threads Executed'), h.index('# Samples')
    data = [r for r in b['rows'] if r[isamp].isdigit()]
    tot, tots = sum(int(r[ie]) for r in data), max(1, sum(int(r[isamp]) for r in data))
    print('=' * 20, 'source of launch', bi, 'instr', tot, 'samples', tots)
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:topn]:
        print(f"{int(r[isamp]) / tots * 100:5.1f}% ex={r[ie]:>11s} thr={r[it]:>3s} {r[ia][:100]}")
