"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of scripts/ncu_pass.py as a markdown table.
usage: python scripts/launch_summary.py launches.csv [setup_launches] > profiles/<name>.md"""
import collections, csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3))          # ns -> us
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for name, us in rows:
    key = re.sub(r"\(.*", "", name)
    key = re.sub(r"^(void )?(ja::)?", "", key)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print("%d launches in the pass, %.1f ms of kernel time in total.\n" % (len(rows), tot / 1e3))
print("| kernel | launches | total ms | share | us / launch |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if v[1] / tot < 0.004:
        continue
    print("| `%s` | %d | %.2f | %.1f %% | %.1f |" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / v[0]))
rk = sum(v[1] for k, v in agg.items() if k.startswith("k_round"))
print("\nRound kernels (`k_round_*`, the `sumcheck_fused` class of bench.py) = %.1f %% of the kernel time of the pass." % (100 * rk / tot))
