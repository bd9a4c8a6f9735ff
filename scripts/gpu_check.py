"""Quick on-GPU sanity run (development aid): parity vs the oracle + timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from multiviewhmr_b200 import synthetic as syn, aggregation as agg

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

def time_fn(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]

dev = torch.device('cuda:0')
print(torch.cuda.get_device_name(0))
w = syn.CONFIGS['cfg1']
f, P, cv, c = syn.make_inputs(w)
for m in ['sum', 'mean', 'max', 'softmax']:
    o = agg.unprojection(f.to(dev), P.to(dev), cv.to(dev), m).cpu().numpy()
    r = oracle.unprojection(f, P, cv, m)
    print('cfg1', m, 'rel', rel(o, r), 'maxabs', float(np.abs(o - r).max()), 'biteq', float((o == r).mean()))
for name in ['unproj_ragged', 'unproj_edge', 'unproj_bf16']:
    z = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', name + '.npz'))
    for m in ['sum', 'mean', 'max', 'softmax']:
        if 'out_' + m not in z: continue
        ft = torch.from_numpy(z['features']).to(dev)
        o = agg.unprojection(ft, torch.from_numpy(z['proj']).to(dev), torch.from_numpy(z['coord_volumes']).to(dev), m).cpu().numpy()
        print(name, m, 'rel', rel(o, z['out_' + m]), 'biteq', float((o == z['out_' + m]).mean()))
        if name == 'unproj_bf16':
            o = agg.unprojection(ft.bfloat16(), torch.from_numpy(z['proj']).to(dev), torch.from_numpy(z['coord_volumes']).to(dev), m).cpu().numpy()
            print(name, m, '(bf16 storage) rel', rel(o, z['out_' + m]), 'biteq', float((o == z['out_' + m]).mean()))

tiles = [None, '32', '16']
for cfgname in ['cfg2', 'cfg3', 'cfg4', 'cfg5']:
    w = syn.CONFIGS[cfgname]
    f, P, cv, c = syn.make_inputs(w)
    fd = f.to(dev); Pd = P.to(dev); cvd = cv.to(dev)
    if w.dtype == 'bf16': fd = fd.bfloat16()
    out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
    ab = w.algorithmic_bytes()
    for tile in tiles:
        if tile is None: os.environ.pop('MVHMR_LZ', None)
        else: os.environ['MVHMR_LZ'] = tile
        tmin, tmed = time_fn(lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out))
        print(f'{cfgname} tile={tile}: min {tmin*1e3:.1f} us med {tmed*1e3:.1f} us  {w.vcv/tmin/1e6:.0f} Gvcv/s  roofline {ab/tmin/1e6/6543.1*100:.1f}%', flush=True)
    os.environ.pop('MVHMR_LZ', None)
    packed = agg.pack_features(fd)
    tmin, tmed = time_fn(lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, packed=packed))
    print(f'{cfgname} prepacked: min {tmin*1e3:.1f} us')
    tmin, tmed = time_fn(lambda: agg.pack_features(fd))
    print(f'{cfgname} pack only: min {tmin*1e3:.1f} us')
    if cfgname in ('cfg2', 'cfg3'):
        t = time.time(); r = oracle.unprojection(f, P, cv, w.method); print('oracle s', time.time() - t)
        o = out.cpu().numpy()
        print(cfgname, 'rel vs oracle', rel(o, r), 'maxabs', float(np.abs(o - r).max()))
    if w.joints:
        vol = out[:, :w.joints].contiguous()
        sa = agg.soft_argmax_3d(vol, cvd)
        tr = oracle.soft_argmax_3d(vol.cpu(), cv)
        print('softargmax maxabs', float(np.abs(sa.cpu().numpy() - tr).max()), 'max|coord|', float(cv.abs().max()))
        tmin, tmed = time_fn(lambda: agg.soft_argmax_3d(vol, cvd))
        print(f'softargmax: min {tmin*1e3:.1f} us  roofline {w.soft_argmax_bytes()/tmin/1e6/6543.1*100:.1f}%')
