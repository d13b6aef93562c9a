"""Per-kernel breakdown of ONE context-encoder forward from an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L.csv python tools/time_context.py 256 1 --no-eager
   python tools/ctx_breakdown.py L.csv 256"""
import csv, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
L = list(csv.DictReader(lines))
names = [(x['Kernel Name'], float(x['Metric Value'].replace(',', '')), x['Grid Size']) for x in L]
last = [i for i, n in enumerate(names) if 'head' in n[0]][-1]
start = [i for i, n in enumerate(names[:last]) if 'image_to_nhwc' in n[0]][-1]
tot = sum(v for n, v, g in names[start:last + 1])
# (output size, Cout, executed K) in execution order: stem, layer1 x4, then (conv1 s2, downsample, conv2, conv1, conv2) x3
specs = [(112, 64, 28 * 64)] + [(56, 64, 576)] * 4
for oh, co, ci in ((28, 128, 64), (14, 256, 128), (7, 512, 256)):
    specs += [(oh, co, 9 * ci), (oh, co, ci), (oh, co, 9 * co), (oh, co, 9 * co), (oh, co, 9 * co)]
ci = 0
for n, v, g in names[start:last + 1]:
    extra = ""
    if 'conv' in n:
        oh, co, k = specs[ci]; ci += 1
        fl = 2.0 * B * oh * oh * co * k
        extra = "  out %3dx%-3d N=%-3d K=%-4d %6.1f GFLOP %6.0f TFLOP/s" % (oh, oh, co, k, fl / 1e9, fl / (v * 1e-9) / 1e12)
    print("%-36s %9.1f us %5.1f%% %s" % (n.replace('<unnamed>::', '')[:36], v / 1e3, 100 * v / tot, extra))
print("total %.1f us for %d agents (serialised, cold cache: shares, not absolutes)" % (tot / 1e3, B))
