"""Turn the gpurun_out/ captures of a round into the tracked summaries under profiles/."""
import csv, json, collections, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
rep = os.path.join(ROOT, 'gpurun_out', 'prof_%s_final_cfg2.ncu-rep' % tag)
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
open('/tmp/raw_final.csv', 'w').write(raw)
summ = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_keys.py'), '/tmp/raw_final.csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(open('/tmp/raw_final.csv'))); d = dict(zip(rr[0], rr[2]))
st = [(float(d[h]), h) for h in rr[0] if 'issue_stalled' in h and 'ratio' in h and d[h] not in ('', 'n/a')]
with open(os.path.join(ROOT, 'profiles', '%s_unproject_cfg2_ncu_full_summary.txt' % tag), 'w') as f:
    f.write('ncu --set full --clock-control none --import-source on -k regex:unproject_kernel -s 2 -c 1 python scripts/prof_run.py cfg2 3\n')
    f.write(summ.replace('-- stalls (warps per issue-active cycle)\n', ''))
    f.write('-- warp stall reasons (warps per issue-active cycle)\n')
    for v, h in sorted(st, reverse=True)[:10]:
        f.write('%8.3f %s\n' % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    for k in ['l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
              'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']:
        if k in d: f.write('%s = %s\n' % (k, d[k]))
rd = float(d['dram__bytes_read.sum']) * 1e6; wr = float(d['dram__bytes_write.sum']) * 1e6
json.dump({"cfg2": int(rd + wr), "_note": "dram__bytes_read.sum + dram__bytes_write.sum of one unproject_kernel launch at cfg2, ncu --set full (profiles/%s_unproject_cfg2_ncu_full_summary.txt); the 331.4 MB algorithmic bytes also count the NCHW feature read done by pack_kernel" % tag,
           "cfg2_read": int(rd), "cfg2_write": int(wr)}, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)
src = os.path.join(ROOT, 'gpurun_out', 'launches_%s.csv' % tag)
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[ix['Metric Value']].replace(',', ''))
    except Exception: continue
    a = agg.setdefault(r[ix['Kernel Name']][:60], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(ROOT, 'profiles', '%s_bench_launch_shares.txt' % tag), 'w') as f:
    f.write('ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline\n(cold-cache, serialised launches: shares, not absolutes)\n')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%-62s n=%4d total %10.1f us  avg %8.1f us  share %5.1f%%\n' % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot))
import shutil
shutil.copy(src, os.path.join(ROOT, 'profiles', '%s_bench_launches.csv' % tag))
line = open(os.path.join(ROOT, 'gpurun_out', 'bench_final.log')).read().strip().splitlines()[-1]
open(os.path.join(ROOT, 'profiles', '%s_bench_line.json' % tag), 'w').write(line + '\n')
print(open(os.path.join(ROOT, 'profiles', '%s_bench_launch_shares.txt' % tag)).read())
print(open(os.path.join(ROOT, 'profiles', '%s_unproject_cfg2_ncu_full_summary.txt' % tag)).read())
dd = json.loads(line); print(dd['value'], dd['ms_per_step'], dd['roofline']['frac'], dd['roofline']['kernel_ms'], dd['e2e']['value'])
