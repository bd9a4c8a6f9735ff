"""Static SASS mnemonic counts of the kernels the bench and the A/B runs launch -> profiles/r2_sass_mnemonics.txt.
usage: python scripts/sass_table.py   (needs cuobjdump + c++filt; no GPU)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'multiviewhmr_b200', 'lib', 'libmvhmr_b200.so')
COLS = ['FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'FMNMX3', 'FMNMX', 'MUFU', 'LDG', 'LDS', 'STS', 'STG', 'TEX', 'UBLKCP', 'SYNCS', 'REDUX', 'REDG', 'LDL', 'STL', 'BAR']
PICK = ['unproject_kernel<8, true, false, false, 3, 7, 0>', 'unproject_kernel<4, true, true, false, 3, 7, 0>',
        'unproject_kernel<4, true, true, true, 3, 6, 0>', 'unproject_kernel<4, true, true, false, 3, 7, 1>',
        'unproject_kernel<4, true, true, false, 3, 7, 2>', 'unproject_kernel<4, true, true, true, 3, 6, 3>',
        'unproject_kernel<4, true, true, false, 3, 7, 3>',
        'unproject_staged_kernel<8, false, 3, 8>', 'unproject_staged_kernel<4, false, 3, 8>',
        'unproject_tex_kernel<4, 3, true, true>', 'unproject_tex_kernel<8, 3, true, true>', 'tex_pack_vec_kernel<true>',
        'pack_kernel<false>', 'soft_argmax_partials_kernel<true, false>', 'soft_argmax_finalize_kernel',
        'unproject_backward_packed_kernel<false, 0, true>', 'unproject_backward_packed_kernel<false, 3, false>']
sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
names = re.findall(r'Function : (\S+)', sass)
dem = dict(zip(names, subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()))
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = per.setdefault(dem[m.group(1)].replace('mvhmr::', '').replace('void ', ''), collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m and cur is not None:
        op = m.group(1)
        cur[op] += 1
        cur['total'] += 1
out = ['cuobjdump -sass multiviewhmr_b200/lib/libmvhmr_b200.so: static instruction counts of the kernels the bench and the A/B runs launch  (scripts/sass_table.py)',
       '(FFMA2/FMUL2/FADD2 = packed fp32x2 math; UBLKCP = cp.async.bulk through the TMA engine; SYNCS = mbarrier; TEX = texture fetch; LDL/STL = spills;',
       ' OUT = 3 (last template argument of unproject_kernel) = the fused unproject + aggregate + soft-argmax kernel)', '',
       '%-62s %6s ' % ('kernel', 'total') + ' '.join('%6s' % c for c in COLS)]
for want in PICK:
    for name, cnt in per.items():
        if name.startswith(want):
            out.append('%-62s %6d ' % (want, cnt['total']) + ' '.join('%6d' % cnt[c] for c in COLS))
            break
    else:
        out.append('%-62s (not in the library)' % want)
tot = collections.Counter()
for cnt in per.values():
    tot.update(cnt)
out += ['', 'whole library (%d kernels): ' % len(per) + ', '.join('%s %d' % (c, tot[c]) for c in COLS)]
open(os.path.join(ROOT, 'profiles', 'r2_sass_mnemonics.txt'), 'w').write('\n'.join(out) + '\n')
print('\n'.join(out))
