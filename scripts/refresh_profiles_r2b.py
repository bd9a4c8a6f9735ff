"""Second batch of round-2 captures (gpurun_out/*.ncu-rep, bench lines, launch list) -> tracked summaries under profiles/."""
import collections, csv, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
KEYS = os.path.join(ROOT, 'scripts', 'ncu_keys.py')
EXTRA = ['l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active',
         'l1tex__t_sector_pipe_tex_mem_texture_op_tex_hit_rate.pct', 'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum']


def summary(rep_name, cmd, dst, note=''):
    rep = os.path.join(GO, rep_name)
    if not os.path.exists(rep):
        print('missing', rep)
        return
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    with open(os.path.join(PR, dst), 'w') as f:
        f.write(cmd + '\n' + note)
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            one = '/tmp/one.csv'
            with open(one, 'w') as fh:
                w = csv.writer(fh); w.writerow(hdr); w.writerow(rows[1]); w.writerow(row)
            s = subprocess.run([sys.executable, KEYS, one], capture_output=True, text=True).stdout
            f.write(s.replace('-- stalls (warps per issue-active cycle)\n', ''))
            st = [(float(d[h]), h) for h in hdr if 'issue_stalled' in h and 'ratio' in h and d[h] not in ('', 'n/a')]
            f.write('-- warp stall reasons (warps per issue-active cycle)\n')
            for v, h in sorted(st, reverse=True)[:8]:
                f.write('%8.3f %s\n' % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
            for k in EXTRA:
                if k in d and d[k] not in ('', 'n/a'):
                    f.write('%s = %s\n' % (k, d[k]))
            f.write('\n')
    print(open(os.path.join(PR, dst)).read())
    d = dict(zip(hdr, rows[-1])); d['__units__'] = dict(zip(hdr, rows[1]))
    return d


def launch_shares(src_name, dst_csv, dst_txt, header):
    src = os.path.join(GO, src_name)
    if not os.path.exists(src):
        print('missing', src)
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[ix['Metric Value']].replace(',', ''))
        except Exception:
            continue
        a = agg.setdefault(r[ix['Kernel Name']][:72], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(PR, dst_txt), 'w') as f:
        f.write(header)
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-74s n=%4d total %10.1f us  avg %8.1f us  share %5.1f%%\n' % (k, a[0], a[1] / 1e3, a[1] / a[0] / 1e3, 100 * a[1] / tot))
    shutil.copy(src, os.path.join(PR, dst_csv))
    print(open(os.path.join(PR, dst_txt)).read())


summary('prof_r2_fused_cfg3.ncu-rep',
        'ncu --set full --clock-control none --import-source on -k regex:unproject_kernel -s 4 -c 2 python scripts/prof_fused.py   (cfg3: bf16 maps, 17 joints)',
        'r2_fused_softargmax_cfg3_ncu_full_summary.txt',
        '(the fused unproject + aggregate + soft-argmax kernel, OUT = 3: first launch with the volume stored, second joints only — out == NULL)\n')
summary('prof_r2_tex_v1_cfg3.ncu-rep',
        'ncu --set full --clock-control none --import-source on -k regex:unproject_tex -s 2 -c 1 python scripts/prof_fast.py cfg3',
        'r2_tex_path_v1_cfg3_ncu_full_summary.txt',
        "(precision='fast', the committed texture-path kernel: plain-FMA position, 16 z x 2 x lane mapping; XU = MUFU is the busiest pipe)\n")
d5 = summary('prof_r2b_cfg5_full.ncu-rep',
        'ncu --set full --clock-control none --import-source on -k regex:unproject_kernel -s 2 -c 1 python scripts/prof_run.py cfg5 3   (B = 64: the headline launch)',
        'r2_unproject_cfg5_v2_ncu_full_summary.txt',
        '(the headline kernel after the output tile moved into the records and the CTA chunks are dealt dynamically; the earlier state is r2_unproject_cfg5_ncu_full_summary.txt)\n')
launch_shares('r2b_bench_launches.csv', 'r2_bench_launches.csv', 'r2_bench_launch_shares.txt',
              'ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline\n'
              '(cold-cache, serialised launches: shares, not absolutes; the first 600 launches cover the headline loop, both e2e legs and the start of the per-config block)\n')
for src, dst in (('r2b_bench_final.json', 'r2_bench_line.json'), ('r2b_bench_n2.json', 'r2_bench_line_n2.json'), ('r2b_bench_n4.json', 'r2_bench_line_n4.json'), ('r2b_bench_n8.json', 'r2_bench_line_n8.json')):
    p = os.path.join(GO, src)
    if os.path.exists(p):
        open(os.path.join(PR, dst), 'w').write(open(p).read().strip().splitlines()[-1] + '\n')

if d5:
    import json
    scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}
    u = d5['__units__']
    rd = float(d5['dram__bytes_read.sum']) * scale[u['dram__bytes_read.sum']]
    wr = float(d5['dram__bytes_write.sum']) * scale[u['dram__bytes_write.sum']]
    tp = os.path.join(PR, 'traffic.json')
    traffic = json.load(open(tp))
    traffic['cfg5'] = int(rd + wr); traffic['cfg5_read'] = int(rd); traffic['cfg5_write'] = int(wr)
    json.dump(traffic, open(tp, 'w'), indent=1)
    print('traffic cfg5', traffic['cfg5'])
