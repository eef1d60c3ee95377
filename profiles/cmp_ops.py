"""Compare two per-op profiles written by quick_perf.py (gpurun_out/profile_ops.txt copies)."""
import sys, collections
def load(p):
    d = {}
    for l in open(p):
        a = l.split(); d[a[0]] = (a[1], float(a[2]))
    return d
a, b = load(sys.argv[1]), load(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
print('total', round(sum(v[1] for v in a.values()), 3), round(sum(v[1] for v in b.values()), 3))
for d in (a, b):
    agg = collections.defaultdict(float)
    for k, v in d.items(): agg[v[0]] += v[1]
    print({k: round(v, 3) for k, v in agg.items()})
for k in a:
    if k in b and abs(a[k][1] - b[k][1]) > thr: print(f"{k:36s} {a[k][1]:.4f} {b[k][1]:.4f}")
