"""Summarise an .ncu-rep: raw metrics of each profiled kernel + hottest source lines (developer tool)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum']
for r in rows[2:]:
    if pat and pat not in r[hdr.index('Kernel Name')]:
        continue
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_warp_active.pct') and 'not_issued' not in h:
            try:
                stalls.append((float(r[i]), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_warp_active.pct', '')))
            except ValueError:
                pass
    print("stalls:", ", ".join(f"{n}={v:.1f}" for v, n in sorted(stalls, reverse=True)[:8]))
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs, cur = [], None
for r in rows:
    if r and r[0] == 'Function Name':
        cur = {'name': r[1], 'lines': []}; secs.append(cur); continue
    if cur is None or not r: continue
    if r[0] in ('File Path', 'Line No'):
        if r[0] == 'Line No': cur['hdr'] = r
        continue
    if r[0] != '':
        try:
            h_ = cur['hdr']; int(r[h_.index('Instructions Executed')]); int(r[h_.index('# Samples')]); int(r[h_.index('Thread Instructions Executed')]); int(r[0])
            cur['lines'].append(r)
        except Exception:
            pass
for s in secs:
    if pat and pat not in s['name']: continue
    if 'hdr' not in s or not s['lines']: continue
    h = s['hdr']; ie = h.index('Instructions Executed'); it = h.index('Thread Instructions Executed'); ss = h.index('# Samples')
    totw = sum(int(l[ie]) for l in s['lines']) or 1; tots = sum(int(l[ss]) for l in s['lines']) or 1
    if totw < 1000: continue
    print(s['name'][:90], 'warp-inst', totw, 'samples', tots)
    top = sorted(s['lines'], key=lambda l: -(int(l[ss]) / tots + int(l[ie]) / totw))[:int(sys.argv[3]) if len(sys.argv) > 3 else 22]
    for l in sorted(top, key=lambda l: int(l[0])):
        print(f"  {l[0]:>4} inst={int(l[ie])/totw*100:5.1f}% thr/inst={int(l[it])/max(int(l[ie]),1):5.1f} samp={int(l[ss])/tots*100:5.1f}%  {l[1].strip()[:105]}")
