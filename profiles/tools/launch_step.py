"""From an ncu launch list (gpu__time_duration.sum csv) covering more than one eager step, cut out exactly ONE step (between
two consecutive optimizer launches, the last kernels of a step) and print / write its per-kernel summary."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
def us(r):
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    return v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
marks = [i for i, r in enumerate(rows) if 'multi_tensor_apply' in r['Kernel Name']]
# the SGD step launches its multi-tensor kernels back to back: a step ends at the last one of a run
ends = [m for j, m in enumerate(marks) if j + 1 == len(marks) or marks[j + 1] != m + 1]
assert len(ends) >= 2, "need at least one full step between two optimizer launches"
a, b = ends[-2] + 1, ends[-1] + 1
step = rows[a:b]
if len(sys.argv) > 2:
    with open(sys.argv[2], 'w') as f:
        w = csv.writer(f); w.writerow(['index', 'kernel', 'grid', 'block', 'duration_us'])
        for i, r in enumerate(step):
            w.writerow([i, r['Kernel Name'], r.get('Grid Size', ''), r.get('Block Size', ''), f"{us(r):.2f}"])
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    k = r['Kernel Name'].replace('void ', '').replace('jmt::', '')[:50]
    agg[k][0] += 1; agg[k][1] += us(r)
tot = sum(v[1] for v in agg.values())
print(len(step), 'launches', round(tot, 1), 'us total (one eager step, cold-cache serialised ncu times)')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} {v[0]:5d} {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0]:7.1f}")
