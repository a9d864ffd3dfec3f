"""Summarise an ncu --set full report: key metrics + top stall locations.  python tools/ncu_summary.py file.ncu-rep"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in want:
        if k in d:
            print(f"{k} = {d[k]} {units[hdr.index(k)]}")
    for k in hdr:
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio'):
            try:
                v = float(d[k])
            except ValueError:
                continue
            if v > 0.3:
                print(f"  stall {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}: {v:.2f}")
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
i_src, i_samp, i_inst = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
seen, data = set(), []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in seen: continue
    seen.add(r[0])
    try: data.append((int(r[i_samp]), int(r[i_inst]), r[i_src].strip()))
    except ValueError: pass
tot = sum(d[0] for d in data) or 1
print(f"SASS lines {len(data)}, samples {tot}, warp-instructions {sum(d[1] for d in data)}")
for s_, n, t in sorted(data, key=lambda d: -d[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"{100*s_/tot:5.1f}%  inst={n:9d}  {t[:100]}")
