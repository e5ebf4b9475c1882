"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, stall mix, hottest SASS."""
import csv, subprocess, sys, io

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_static', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__waves_per_multiprocessor',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum',
        'smsp__inst_executed_op_global_ld.sum', 'smsp__inst_executed_op_global_st.sum']
for r in rows[2:]:
    print('=' * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i][:80]} {units[i]}")
    st = []
    for i, h in enumerate(hdr):
        if 'average_warps_issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            st.append((float(r[i]), h.split('issue_stalled_')[1].split('_per_issue')[0]))
    print("stall (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
    pipes = []
    for i, h in enumerate(hdr):
        if h.startswith('sm__inst_executed_pipe_') and h.endswith('.avg.pct_of_peak_sustained_active'):
            pipes.append((float(r[i]), h[len('sm__inst_executed_pipe_'):].split('.')[0]))
    print("pipes (% of peak):", ", ".join(f"{n}={v:.1f}" for v, n in sorted(pipes, reverse=True)[:8]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
print('=' * 100)
print(rows[0][:2])
hdr = rows[1]
iS, iSrc, iE = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
seen = set()
for k, r in enumerate(rows[2:]):
    try:
        key = r[0]
        if key in seen:
            continue
        seen.add(key)
        data.append((int(r[iS]), k, r))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
print("instructions:", len(data), "samples:", tot)
opc = {}
for n, k, r in data:
    op = r[iSrc].split()[0] if r[iSrc].split() else '?'
    if op.startswith('@'):
        op = r[iSrc].split()[1]
    op = op.split('.')[0]
    e = int(r[iE] or 0)
    a = opc.setdefault(op, [0, 0])
    a[0] += e
    a[1] += n
tote = sum(v[0] for v in opc.values()) or 1
print("opcode mix (executed %, stall-sample %):")
for op, (e, n) in sorted(opc.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"   {op:10s} exec {100*e/tote:5.1f}%  samples {100*n/tot:5.1f}%")
print("hottest instructions:")
for n, k, r in sorted(data, key=lambda d: -d[0])[:25]:
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"  {100*n/tot:4.1f}%  {r[iSrc][:70]:70s} {top}")
