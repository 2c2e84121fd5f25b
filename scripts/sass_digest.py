"""Instruction histogram per kernel of the in-tree library (cuobjdump -sass): the Blackwell-specific mnemonics that prove the tcgen05 /
TMEM / bulk-copy paths (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, SYNCS = mbarrier, UTCBAR =
tcgen05.commit, REDUX = warp reductions, RED = fire-and-forget atomics). Usage: python scripts/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'alpha-zero-general_b200', 'csrc', 'libazg_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'SYNCS', 'REDUX', 'HMMA', 'DFMA', 'DADD', 'DMUL', 'FFMA', 'FFMA2', 'LDG', 'STG', 'LDS', 'STS', 'ATOMG', 'RED', 'BAR', 'SHFL', 'VOTE', 'MUFU']
cur = None; hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1); hist[cur]['_total'] += 1
        for k in KEYS:
            if op == k or op.startswith(k + '.') or (k in ('UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'SYNCS', 'REDUX', 'ATOMG', 'RED') and op.startswith(k)):
                hist[cur][k] += 1
def demangle(n):
    try:
        return subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip().split('(')[0][:70] or n[:70]
    except Exception:
        return n[:70]
print('# cuobjdump -sass alpha-zero-general_b200/csrc/libazg_b200.so: SASS instruction counts per kernel (static code, not executed counts)')
print('%-72s %8s  %s' % ('kernel', 'instr', 'mnemonic counts'))
for fn, h in sorted(hist.items(), key=lambda kv: -kv[1]['_total']):
    if h['_total'] < 200: continue
    print('%-72s %8d  %s' % (demangle(fn), h['_total'], ' '.join(f'{k}={h[k]}' for k in KEYS if h[k])))
