"""Per-barrier-region view of an exported ncu source page (development aid):

    ncu -i rep.ncu-rep --page source --csv > src.csv
    python scripts/ncu_regions.py src.csv [positions] [warps_per_cta]

Regions are the SASS ranges between BAR / EXIT instructions.  For each one:
share of the stall samples, share of the executed warp instructions, the five
largest stall reasons (% of the region's samples) and warp-instructions per
position and warp for the memory / math opcodes."""
import csv
import sys


def main(path, positions=20000, warps=16):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    S, I, SRC = ci['# Samples'], ci['Instructions Executed'], ci['Source']
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[S]) for r in data)
    toti = sum(int(r[I]) for r in data)

    def op(r):
        t = r[SRC].split()
        return t[0] if not t[0].startswith('@') else t[1]

    regions, cur = [], []
    for idx, r in enumerate(data):
        cur.append((idx, r))
        o = op(r)
        if o.startswith('BAR') or o.startswith('EXIT'):
            regions.append(cur)
            cur = []
    if cur:
        regions.append(cur)
    print('total samples', tot, 'warp inst', toti, 'regions', len(regions))
    key = ['LDG', 'STG', 'LDS', 'STS', 'REDG', 'LDTM', 'STTM', 'FFMA', 'FMUL', 'FADD',
           'FADD2', 'MUFU', 'CCTL', 'BAR']
    for reg in regions:
        s = sum(int(r[S]) for _, r in reg)
        ie = sum(int(r[I]) for _, r in reg)
        if s < 0.003 * tot:
            continue
        st = {k: sum(int(r[ci[k]] or 0) for _, r in reg) for k in stalls}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
        ops = {}
        for _, r in reg:
            o = op(r).split('.')[0]
            ops[o] = ops.get(o, 0) + int(r[I])
        print(f"[{reg[0][0]:5d}-{reg[-1][0]:5d}] samp {100 * s / tot:5.1f}% inst {100 * ie / toti:5.1f}% "
              f"ratio {(s / tot) / (ie / toti + 1e-9):4.2f} | "
              + ' '.join(f"{k[6:]}={100 * v / s:.0f}" for k, v in top) + ' | '
              + ' '.join(f"{k}:{ops.get(k, 0) / positions / warps:.0f}" for k in key if ops.get(k, 0)))


if __name__ == '__main__':
    main(sys.argv[1], *(int(a) for a in sys.argv[2:]))
