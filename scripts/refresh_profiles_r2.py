"""Turn the round-2 gpurun_out/ captures into the tracked summaries under profiles/."""
import csv, json, collections, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
KEYS = os.path.join(ROOT, 'scripts', 'ncu_keys.py')


def raw(rep):
    out = '/tmp/%s.csv' % os.path.basename(rep)
    open(out, 'w').write(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout)
    return out


def summary(rep_name, cmd, dst, extra_note=''):
    rep = os.path.join(GO, rep_name)
    if not os.path.exists(rep):
        print('missing', rep); return None
    r = raw(rep)
    rows = list(csv.reader(open(r)))
    hdr = rows[0]
    with open(os.path.join(PR, dst), 'w') as f:
        f.write(cmd + '\n' + extra_note)
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            one = '/tmp/one.csv'
            w = csv.writer(open(one, 'w')); w.writerow(hdr); w.writerow(rows[1]); w.writerow(row)
            del w
            s = subprocess.run([sys.executable, KEYS, one], capture_output=True, text=True).stdout
            f.write(s.replace('-- stalls (warps per issue-active cycle)\n', ''))
            st = [(float(d[h]), h) for h in hdr if 'issue_stalled' in h and 'ratio' in h and d[h] not in ('', 'n/a')]
            f.write('-- warp stall reasons (warps per issue-active cycle)\n')
            for v, h in sorted(st, reverse=True)[:8]:
                f.write('%8.3f %s\n' % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            for k in ['l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_pipe_tex_mem_texture_op_tex_hit_rate.pct',
                      'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
                      'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum']:
                if k in d and d[k] not in ('', 'n/a'): f.write('%s = %s\n' % (k, d[k]))
            f.write('\n')
    rows = list(csv.reader(open(r)))
    d = dict(zip(rows[0], rows[2]))
    d['__units__'] = dict(zip(rows[0], rows[1]))
    return d


def dram_bytes(_unused, d):
    scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
    u = d['__units__']
    return (float(d['dram__bytes_read.sum']) * scale[u['dram__bytes_read.sum']],
            float(d['dram__bytes_write.sum']) * scale[u['dram__bytes_write.sum']])


d5 = summary('prof_r2_cfg5_full.ncu-rep', 'ncu --set full --clock-control none --import-source on -k regex:unproject_kernel -s 2 -c 1 python scripts/prof_run.py cfg5 3   (B = 64: the headline launch)', 'r2_unproject_cfg5_ncu_full_summary.txt')
d4 = summary('prof_r2_cfg4_full.ncu-rep', 'ncu --set full --clock-control none --import-source on -k regex:unproject_kernel -s 2 -c 1 python scripts/prof_run.py cfg4 3', 'r2_unproject_cfg4_ncu_full_summary.txt')
summary('prof_r2_staged_v0_cfg2.ncu-rep', 'MVHMR_PATH=staged MVHMR_STAGED_CTAS=2 ncu --set full ... -k regex:unproject_staged python scripts/prof_run.py cfg2 3', 'r2_staged_cfg2_ncu_full_summary.txt',
        '(the shared-memory-staged kernel, first working version: opt-in, lost the A/B — DESIGN.md section 4)\n')
summary('prof_r2_staged_v0_cfg5.ncu-rep', 'MVHMR_PATH=staged MVHMR_STAGED_CTAS=2 ncu --set full ... -k regex:unproject_staged python scripts/prof_run.py cfg5 3 8   (B = 8)', 'r2_staged_cfg5_ncu_full_summary.txt',
        '(most bricks overflow the 97-pixel patches at two CTAs per SM and read through the global fallback)\n')
summary('prof_r2_tex_v0_cfg3.ncu-rep', 'ncu --set full ... -k regex:"unproject_tex|tex_pack" -s 4 -c 2 python scripts/prof_fast.py cfg3', 'r2_tex_path_cfg3_ncu_full_summary.txt',
        "(precision='fast', FIRST version of the texture-path kernel: 178 us, issue-bound at 132 M warp instructions; the committed kernel drops the per-view / per-channel guards: 94 instructions per channel quad, 44 registers)\n")

traffic = json.load(open(os.path.join(PR, 'traffic.json')))
for name, d in (('cfg5', d5), ('cfg4', d4)):
    if d:
        rd, wr = dram_bytes(None, d)
        traffic[name] = int(rd + wr); traffic[name + '_read'] = int(rd); traffic[name + '_write'] = int(wr)
traffic['_note'] = ("dram__bytes_read.sum + dram__bytes_write.sum of one unproject_kernel launch, ncu --set full "
                    "(profiles/r1_unproject_cfg2_ncu_full_summary.txt, profiles/r2_unproject_cfg{4,5}_ncu_full_summary.txt); "
                    "the algorithmic bytes also count the NCHW feature read done by pack_kernel")
json.dump(traffic, open(os.path.join(PR, 'traffic.json'), 'w'), indent=1)

src = os.path.join(GO, 'r2_bench_launches.csv')
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[ix['Metric Value']].replace(',', ''))
    except Exception: continue
    a = agg.setdefault(r[ix['Kernel Name']][:72], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(PR, 'r2_bench_launch_shares.txt'), 'w') as f:
    f.write('ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline\n'
            '(cold-cache, serialised launches: shares, not absolutes; the first 600 launches cover the headline loop, both e2e legs and the start of the per-config block)\n')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%-74s n=%4d total %10.1f us  avg %8.1f us  share %5.1f%%\n' % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot))
shutil.copy(src, os.path.join(PR, 'r2_bench_launches.csv'))
line = open(os.path.join(GO, 'r2_bench_final.log')).read().strip().splitlines()[-1]
open(os.path.join(PR, 'r2_bench_line.json'), 'w').write(line + '\n')
if os.path.exists(os.path.join(GO, 'r2_bench_n2.log')):
    open(os.path.join(PR, 'r2_bench_line_n2.json'), 'w').write(open(os.path.join(GO, 'r2_bench_n2.log')).read().strip().splitlines()[-1] + '\n')
with open(os.path.join(PR, 'r2_microbenchmarks.txt'), 'w') as f:
    for name, title in (('bulk_bench2.log', 'scripts/micro/bulk_bench.cu (first version: per-pixel / per-row cp.async.bulk, cp.async 16 B, ldg+sts; 512 planes = DRAM-resident, 8 = L2-resident)'),
                        ('bulk_bench3.log', 'scripts/micro/bulk_bench.cu (per-row copies issued from one warp vs from one warp per view with an mbarrier per view)'),
                        ('l1_bench.log', 'scripts/micro/l1_bench.cu (L1 cost of a 16-byte-per-lane gather by lines touched; LDS.128 for comparison)'),
                        ('tex_bench.log', 'scripts/micro/tex_bench.cu (hardware bilinear filtering of half4 texels: rate and error)'),
                        ('path1.log', 'scripts/path_bench.py (MVHMR_PATH=gather|staged, first working staged kernel, default knobs)'),
                        ('path2.log', 'scripts/path_bench.py staged only: MVHMR_STAGED_T (threads per voxel) x MVHMR_STAGED_CTAS (CTAs per SM)'),
                        ('var1.log', 'scripts/variant_bench.py: compile-time variants of the gather kernel at V = 8 (MVHMR_TV views in registers, MVHMR_WARPS)'),
                        ('var2.log', 'scripts/variant_bench.py: two voxels in flight per lane group, views fused 1 / 2 / 4 at a time (reverted)')):
        pth = os.path.join(GO, name)
        if os.path.exists(pth):
            f.write('==== %s\n%s\n' % (title, open(pth).read()))
print(open(os.path.join(PR, 'r2_bench_launch_shares.txt')).read())
print(open(os.path.join(PR, 'r2_unproject_cfg5_ncu_full_summary.txt')).read())
