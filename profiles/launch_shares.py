"""Per-kernel shares and DRAM traffic of one NFE from an ncu launch list
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file <csv>).
usage: python profiles/launch_shares.py gpurun_out/launches_v2.csv [traffic.json]"""
import csv, sys, json, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
per = collections.OrderedDict()
for r in rows:
    lid, name, metric, val = r[0], r[4].split("(")[0], r[12], float(r[14])
    per.setdefault(lid, {"name": name})[metric] = val
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    a = agg[d["name"]]
    a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0) / 1e3
    a[2] += d.get("dram__bytes_read.sum", 0); a[3] += d.get("dram__bytes_write.sum", 0)
tot = sum(a[1] for a in agg.values())
print("# one NFE (CIFAR, batch 1024), ncu --clock-control none (cold-cache, serialised: compare SHARES); DRAM bytes summed per kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} launches={a[0]:3d} total_us={a[1]:9.1f} share={100 * a[1] / tot:5.1f}%  dram_read_MB={a[2] / 1e6:9.1f} dram_write_MB={a[3] / 1e6:9.1f}")
print(f"TOTAL launches={sum(a[0] for a in agg.values())} total_us={tot:.1f} dram_GB={sum(a[2] + a[3] for a in agg.values()) / 1e9:.2f}")
if len(sys.argv) > 2:
    conv = [a for k, a in agg.items() if "conv_tc" in k]
    n = sum(a[0] for a in conv)
    json.dump({"dram_bytes_per_launch": sum(a[2] + a[3] for a in conv) / n, "launches": n,
               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the conv_tc*/conv_tc2 launches of one NFE (" + sys.argv[1] + ")"},
              open(sys.argv[2], "w"), indent=1)
