"""Summarise an exported ncu report: key raw metrics + samples per SASS region."""
import csv, re, sys
raw, src = sys.argv[1], sys.argv[2]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 200
rows = list(csv.reader(open(raw)))
d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
for k in ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
          'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
          'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active']:
    print(k, d.get(k))
for k, v in d.items():
    if re.search(r'smsp__average_warps_issue_stalled.*_per_issue_active', k) and float(v[1]) > 0.1:
        print('stall', k.split('stalled_')[1].split('_per')[0], v[1])
rows = list(csv.reader(open(src)))
hdr = rows[1]; data = rows[2:]
ci = {h: i for i, h in enumerate(hdr)}
S = ci['# Samples']; I = ci['Instructions Executed']; SRC = ci['Source']
tot = sum(int(r[S]) for r in data); toti = sum(int(r[I]) for r in data)
print('sass instructions', len(data), 'samples', tot, 'warp inst executed', toti)
for c in range(0, len(data), chunk):
    seg = data[c:c + chunk]
    s = sum(int(r[S]) for r in seg); ie = sum(int(r[I]) for r in seg)
    ops = {}
    for r in seg:
        t = r[SRC].split()
        op = t[0] if not t[0].startswith('@') else t[1]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + int(r[I])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    print(f'{c:6d} samples {100*s/tot:5.1f}% inst {100*ie/toti:5.1f}%', [(k, round(v / 1e6, 1)) for k, v in top])
