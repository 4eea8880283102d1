"""ncu launch-list CSV -> markdown table of one training step (between the AdamW launches of consecutive steps)."""
import collections, csv, re, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = []
with open(src) as f:
    lines = [l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    if row.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(row['Metric Value'].replace(',', '')); unit = row['Metric Unit']
        v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
        rows.append((row['Kernel Name'], v))
idx = [i for i, (k, v) in enumerate(rows) if 'adamw' in k]
seg = rows[idx[1] + 1: idx[3] + 1]
def short(k):
    m = re.search(r'(\w+_kernel)', k)
    return m.group(1) if m else ('torch: ' + k[:60])
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in seg:
    agg[short(k)][0] += 1; agg[short(k)][1] += v
tot = sum(v for _, v in agg.values())
out = [f'# {title}', '', f'`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 1 --warmup 1`; one supervised training step',
       f'(config 2: 512x512, N=21, batch 16): {len(seg)} launches, {tot/1e3:.1f} ms of kernel time (cold-cache, serialised: compare shares).', '',
       '| kernel | launches | us | share |', '|---|---|---|---|']
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f'| {k} | {n} | {v:.0f} | {100*v/tot:.1f}% |')
open(dst, 'w').write('\n'.join(out) + '\n')
print(f'{len(seg)} launches {tot/1e3:.1f} ms -> {dst}')
