"""Per-basic-block instruction/sample shares of one kernel from an .ncu-rep captured with --import-source on."""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
thresh = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if 'Address' in r)
ix = {n: i for i, n in enumerate(rows[h])}
data = []
for r in rows[h + 1:]:
    try:
        data.append((r[ix['Source']], int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])))
    except Exception:
        pass
tot = sum(d[1] for d in data)
ss = sum(d[2] for d in data) or 1
print(f"{kern}: {tot} warp instructions, {ss} samples, {len(data)} SASS instructions")
groups, cur = [], None
for i, (src, n, s) in enumerate(data):
    if cur and abs(n - cur['n']) <= 0.02 * max(n, cur['n'], 1):
        cur['cnt'] += 1; cur['inst'] += n; cur['s'] += s; cur['end'] = i
    else:
        cur = {'start': i, 'end': i, 'n': n, 'cnt': 1, 'inst': n, 's': s, 'first': src}
        groups.append(cur)
for g in groups:
    if g['inst'] / tot > thresh or g['s'] / ss > thresh:
        print(f"[{g['start']:4d}-{g['end']:4d}] exec {g['n']/1e6:7.2f}M x {g['cnt']:3d} inst = {100*g['inst']/tot:5.1f}% inst {100*g['s']/ss:5.1f}% samples   {g['first'][:70]}")
if len(sys.argv) > 4:
    a, b = map(int, sys.argv[4].split('-'))
    for i in range(a, b + 1):
        print(f"{i:4d} {data[i][1]/1e6:7.2f}M {100*data[i][2]/ss:5.1f}%  {data[i][0][:100]}")
