"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/launch_summary.py file.csv [launches_per_step [vocoder_launches]]"""
import collections, csv, re, sys
path = sys.argv[1]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 389
nvoc = int(sys.argv[3]) if len(sys.argv) > 3 else 51     # launches of the vocoder at the end of a step
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
vals = [(r['Kernel Name'], float(r['Metric Value']), r['Grid Size']) for r in rows if r['Metric Name'] == 'gpu__time_duration.sum']
step = vals[-per:]
tot = sum(v for _, v, _ in step)
print(f"launches in file {len(vals)}; last step: {per} launches, {tot/1e6:.3f} ms (cold-cache, serialised)")
agg = collections.defaultdict(lambda: [0, 0.0])
for name, v, g in step:
    name = re.sub(r'\((CUtensorMap|const|void|long|float|int).*', '', name)
    agg[name][0] += 1; agg[name][1] += v
for name, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{v/tot*100:6.2f}%  {v/1e6:8.3f} ms  x{c:4d}  {name[:100]}")
voc = step[-nvoc:]
print(f"vocoder (last {nvoc} launches): {sum(v for _, v, _ in voc)/1e6:.3f} ms; acoustic model: {sum(v for _, v, _ in step[:-nvoc])/1e6:.3f} ms")
agg2 = collections.defaultdict(lambda: [0, 0.0])
for name, v, g in voc:
    m = re.search(r'conv_igemm_kernel<\(int\)(\d+), \(int\)(\d+)', name)
    k = m.groups() if m else name
    agg2[k][0] += 1; agg2[k][1] += v
print("vocoder by tile <BN,BK>:", {k: (c, round(v/1e6, 3)) for k, (c, v) in agg2.items()})
