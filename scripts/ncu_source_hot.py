"""Aggregate an ncu --page source csv (cuda,sass view) by source file:line -> samples and instructions. Usage: ncu_source_hot.py rep [topN]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
lines = out.splitlines()
cur_file = None; agg = collections.OrderedDict(); tot_s = 0; tot_i = 0
for row in csv.reader(lines):
    if not row: continue
    if row[0] == 'File Path': cur_file = row[1].split('/')[-1]; continue
    if not row[0].isdigit() or len(row) < 8: continue
    try: s = int(row[-60]) if False else int(row[6]); i = int(row[7])
    except Exception:
        nums = [x for x in row[2:] if x.isdigit()]
        if len(nums) < 4: continue
        s, i = int(nums[2]), int(nums[3])
    a = agg.setdefault((cur_file, int(row[0])), [0, 0, row[1][:120]])
    a[0] += s; a[1] += i; tot_s += s; tot_i += i
print('total samples', tot_s, 'total warp instructions', tot_i)
for (f, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print('%5.1f%% smp %5.1f%% ins  %s:%d  %s' % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), f, ln, src.strip()))
