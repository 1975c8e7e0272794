"""Summarise an .ncu-rep: per-kernel headline metrics and the hottest source lines (by warp stall samples)."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor']
idx = [(k, hdr.index(k)) for k in keys if k in hdr]
names = []
for r in rows[2:]:
    nm = r[hdr.index('Kernel Name')].split('(')[0]
    names.append(nm)
    print(nm)
    for k, i in idx[1:]:
        print("   %-75s %s %s" % (k, r[i], rows[1][i]))
if pat:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # may contain several kernels; take the first table
    h = None
    data = []
    for r in rows:
        if 'Source' in r and '# Samples' in r:
            if h is not None: break
            h = r; continue
        if h and len(r) == len(h): data.append(r)
    isrc, isamp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    tot = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
    tex = sum(int(r[iex]) for r in data if r[iex].isdigit())
    print("total samples", tot, "warp instr", tex, "sass lines", len(data))
    data.sort(key=lambda r: -int(r[isamp]) if r[isamp].isdigit() else 0)
    for r in data[:topn]:
        print("  %6.2f%%  ex=%-10s %s" % (100.0 * int(r[isamp]) / max(tot, 1), r[iex], r[isrc].strip()[:110]))
