"""Summarise an ncu report: per-kernel headline metrics and the hottest SASS instructions of one launch (stall samples).
    python tools/ncu_hot.py gpurun_out/x.ncu-rep [launch_index]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg']
for r in rows[2:]:
    print('----')
    for w in want:
        if w in hdr:
            print('  ', w, '=', r[hdr.index(w)], units[hdr.index(w)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
si, so, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
d = []
for r in rows[2:]:
    if len(r) > si and r[si].isdigit() and (not d or d[-1][0] != r[0]):
        d.append(r)
half = len(d) // 2 if len(d) > 1 and d[0][so] == d[len(d) // 2][so] else len(d)
d = d[:half]
tot = sum(int(r[si]) for r in d)
print('total samples', tot, 'instructions', len(d))
for i, r in enumerate(d):
    if 'TRYWAIT' in r[so] or 'BAR.SYNC' in r[so]:
        m = re.search(r'\+0x([0-9a-f]+)\]', r[so])
        ns = sum(int(x[si]) for x in d[i:i + 3])
        print('  wait @%d exec %s off %s samples %d  %s' % (i, r[ie], m.group(1) if m else '-', ns, r[so][:60]))
print('bins of 100 instructions:', [sum(int(r[si]) for r in d[s:s + 100]) for s in range(0, len(d), 100)])
top = sorted(range(len(d)), key=lambda i: -int(d[i][si]))[:30]
for i in sorted(top):
    print('%5d %6s %9s  %s' % (i, d[i][si], d[i][ie], d[i][so][:110]))
