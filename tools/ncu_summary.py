#!/usr/bin/env python3
"""Summarise an .ncu-rep (run here, no GPU needed): key metrics + dynamic instruction mix + stalls."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for d in data:
    print('===', d[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f'{w:72s} {d[i]:>18s} {units[i]}')
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
            try:
                v = float(d[i])
            except ValueError:
                continue
            if v > 0.15:
                print(f'  stall {h.split("issue_stalled_")[1].split("_per_issue")[0]:28s} {v:6.2f}')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if starts:
    hdr = rows[starts[0]]
    end = starts[1] - 1 if len(starts) > 1 else len(rows)
    iS, iI = hdr.index('Source'), hdr.index('Instructions Executed')
    tot, byop = 0, collections.Counter()
    for r in rows[starts[0] + 1:end]:
        try:
            n = int(r[iI])
        except (ValueError, IndexError):
            continue
        toks = r[iS].split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        byop[op.split('.')[0]] += n
        tot += n
    print('dynamic warp instructions', tot)
    print('  '.join(f'{op}:{100*n/tot:.1f}%' for op, n in byop.most_common(16)))
