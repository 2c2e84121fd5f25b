"""Debug: V80 tcgen05 kernel vs fp32 kernel vs golden; prints max errors and timing."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import azg_b200
from azg_b200.nnet import NNetWrapper
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
game = azg_b200.SplendorGame()
kat = np.load(os.path.join(G, 'splendor_kat.npz'))
for tag in ('rand', 'shipped'):
    z = np.load(os.path.join(G, f'splendor_v80_{tag}.npz')); sd = {k[4:]: z[k] for k in z.files if k.startswith('sd__')}
    os.environ.pop('AZG_V80_KERNEL', None); net_tc = NNetWrapper(game, {'nn_version': 80}, state_dict=sd)
    os.environ['AZG_V80_KERNEL'] = 'fp32'; net_f = NNetWrapper(game, {'nn_version': 80}, state_dict=sd); os.environ.pop('AZG_V80_KERNEL')
    for n in (96,):
        pi, v = net_tc.predict_batch(z['boards'][:n], z['valids'][:n]); pf, vf = net_f.predict_batch(z['boards'][:n], z['valids'][:n])
        print(tag, n, 'tc vs golden: pi %.3e v %.3e | fp32 vs golden: pi %.3e v %.3e | nan %d' % (np.abs(pi - z['pi'][:n]).max(), np.abs(v - z['v'][:n]).max(),
              np.abs(pf - z['pi'][:n]).max(), np.abs(vf - z['v'][:n]).max(), int(np.isnan(pi).sum() + np.isnan(v).sum())), flush=True)
        if np.abs(pi - z['pi'][:n]).max() > 1e-4:
            bad = np.abs(pi - z['pi'][:n]).max(axis=1); print('  per-leaf max err (first 32):', np.array2string(bad[:32], precision=2)); print('  v err:', np.abs(v - z['v'][:n]).max(axis=1)[:32])
    b = np.concatenate([kat['canonical']] * 27)[:16384]; va = np.concatenate([kat['valids']] * 27)[:16384]
    pi, v = net_tc.predict_batch(b, va); pf, vf = net_f.predict_batch(b, va)
    print(tag, 16384, 'tc vs fp32: pi %.3e v %.3e' % (np.abs(pi - pf).max(), np.abs(v - vf).max()), flush=True)

# ---- phase timestamps of CTA 0 (SM clock cycles), AZG_V80_PROF=1 ----
import ctypes as C
os.environ['AZG_V80_PROF'] = '1'
z = np.load(os.path.join(G, 'splendor_v80_rand.npz')); sd = {k[4:]: z[k] for k in z.files if k.startswith('sd__')}
net = NNetWrapper(game, {'nn_version': 80}, state_dict=sd)
b = np.concatenate([kat['canonical']] * 27)[:16384]; va = np.concatenate([kat['valids']] * 27)[:16384]
net.predict_batch(b, va); net.predict_batch(b, va)
out = (C.c_longlong * 64)()
L = net.net._L; L.azg_net_prof.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
print('prof rc', L.azg_net_prof(net.net.h, out))
ts = np.array(list(out), dtype=np.int64); ts = ts[ts != 0]
names = ['tile start', 'input staged']                           # first_layer is folded into the trunk block (no phase of its own)
for b_ in range(3):
    names += [f'b{b_} expand MMA (ch 0-127) done', f'b{b_} depthwise', f'b{b_} fc landed', f'b{b_} SE done', f'b{b_} project MMA done', f'b{b_} project epilogue']
    if b_ == 1: names += ['policy head done']
names += ['tile end']
d = np.diff(ts)
if os.environ.get('ROUND_PROF'):
    print('raw stamp deltas of the first tile:', d[:45].tolist())
for i in range(min(len(d), len(names) - 1)): print('%-26s +%7d cycles' % (names[i + 1], d[i]))
print('tile total', ts[min(len(ts), len(names)) - 1] - ts[0], 'cycles;  stamps', len(ts))
n1 = len(names)
if len(ts) >= 2 * n1:
    d2 = np.diff(ts[n1:2 * n1])
    print('second tile of CTA 0: ' + ', '.join('%s +%d' % (names[i + 1], d2[i]) for i in range(min(2, len(d2)))) + '; total %d cycles' % (ts[2 * n1 - 1] - ts[n1]))

# ---- per-CTA timeline of the last launch (globaltimer ns): entry, prologue done, exit, SM id ----
out2 = (C.c_longlong * 640)()
L.azg_net_prof_ctas.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
if L.azg_net_prof_ctas(net.net.h, out2) == 0:
    a = np.array(list(out2), dtype=np.int64).reshape(160, 4); a = a[a[:, 0] != 0]
    t0 = a[:, 0].min()
    ent, pro, ex = (a[:, 0] - t0) / 1e3, (a[:, 1] - a[:, 0]) / 1e3, (a[:, 2] - t0) / 1e3
    ntile = np.array([len(range(i, 1024, len(a))) for i in range(len(a))])
    body = (a[:, 2] - a[:, 1]) / 1e3
    print('CTAs %d | entry us: min %.1f max %.1f | prologue us: min %.1f med %.1f max %.1f | exit us: min %.1f med %.1f max %.1f' % (
        len(a), ent.min(), ent.max(), pro.min(), np.median(pro), pro.max(), ex.min(), np.median(ex), ex.max()))
    for k in (6, 7):
        m = ntile == k
        if m.any(): print('  CTAs with %d tiles: %d, us per tile: min %.2f med %.2f max %.2f' % (k, m.sum(), (body[m] / k).min(), np.median(body[m] / k), (body[m] / k).max()))
    slow = np.argsort(-ex)[:8]
    print('  slowest CTAs (cta, sm, tiles, exit us, us/tile):', [(int(i), int(a[i, 3]), int(ntile[i]), round(float(ex[i]), 1), round(float(body[i] / ntile[i]), 2)) for i in slow])
