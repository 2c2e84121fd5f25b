"""Debug: V89 tcgen05 kernel vs fp32 kernel vs the reference goldens (max errors) and kernel time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import azg_b200
from azg_b200.nnet import SantoriniNNetWrapper
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
game = azg_b200.SantoriniGame()
for tag in ('rand', 'shipped'):
    z = np.load(os.path.join(G, f'santorini_v89_{tag}.npz')); sd = {k[4:]: z[k] for k in z.files if k.startswith('sd__')}
    os.environ.pop('AZG_V89_KERNEL', None); net_tc = SantoriniNNetWrapper(game, {'nn_version': 89}, state_dict=sd)
    os.environ['AZG_V89_KERNEL'] = 'fp32'; net_f = SantoriniNNetWrapper(game, {'nn_version': 89}, state_dict=sd); os.environ.pop('AZG_V89_KERNEL')
    b, va = z['boards'], z['valids']
    pi, v = net_tc.predict_batch(b, va); pf, vf = net_f.predict_batch(b, va)
    print(tag, len(b), 'tc vs golden: pi %.3e v %.3e | fp32 vs golden: pi %.3e v %.3e | mean signed v err (|tc|-|gold|) %.3e' % (
        np.abs(pi - z['pi']).max(), np.abs(v - z['v']).max(), np.abs(pf - z['pi']).max(), np.abs(vf - z['v']).max(), (np.abs(v) - np.abs(z['v'])).mean()), flush=True)
    for n in (1, 6, 7, 8, 15):
        p2, v2 = net_tc.predict_batch(b[:n], va[:n])
        assert (p2 == pi[:n]).all() and (v2 == v[:n]).all(), n
    bb = np.concatenate([b] * 70)[:4096]; vv = np.concatenate([va] * 70)[:4096]
    for net, name in ((net_tc, 'tcgen05'), (net_f, 'fp32')):
        net.predict_batch(bb, vv); t0 = time.perf_counter(); net.predict_batch(bb, vv); print('  ', name, '4096 leaves incl. copies: %.2f ms' % (1e3 * (time.perf_counter() - t0)))

# ---- phase timestamps of CTA 0 (SM clock cycles), AZG_V89_PROF=1 ----
import ctypes as C
os.environ['AZG_V89_PROF'] = '1'
z = np.load(os.path.join(G, 'santorini_v89_rand.npz')); sd = {k[4:]: z[k] for k in z.files if k.startswith('sd__')}
net = SantoriniNNetWrapper(game, {'nn_version': 89}, state_dict=sd)
bb = np.concatenate([z['boards']] * 70)[:4096]; vv = np.concatenate([z['valids']] * 70)[:4096]
net.predict_batch(bb, vv)
out = (C.c_longlong * 64)()
L = net.net._L; L.azg_net_prof.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
print('prof rc', L.azg_net_prof(net.net.h, out))
ts = np.array(list(out), dtype=np.int64); ts = ts[ts != 0]
d = np.diff(ts)
names = ['first layer'] + sum([[f'conv{c} MMA', f'conv{c} epilogue'] for c in range(10)], []) + ['heads', '(next tile) first layer']
for i in range(min(len(d), len(names))): print('%-24s +%7d cycles' % (names[i], d[i]))
print('tile total', ts[22] - ts[0] if len(ts) > 22 else None, 'cycles; stamps', len(ts))
print('issuer waited for weights (first tile):', int(out[63]), 'cycles')
