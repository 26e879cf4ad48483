"""Summarise `ncu -i <rep> --page raw --csv` exports of single-launch `--set full` captures into a JSON + markdown table.
usage: python scripts/ncu_raw_summary.py <out_prefix> name=raw.csv:pairs:algo_read:algo_write ..."""
import csv
import json
import sys

KEYS = {"kernel": "Kernel Name", "grid": "Grid Size", "block": "Block Size", "time_us": "gpu__time_duration.sum",
        "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "registers": "launch__registers_per_thread",
        "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "warp_insts": "smsp__inst_executed.sum", "l2_hit_pct": "lts__t_sector_hit_rate.pct"}


def scale(v, unit):
    m = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "second": 1e6}
    return m.get(unit, 1)


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for k, name in KEYS.items():
        if name in hdr:
            i = hdr.index(name)
            v = vals[i].replace(",", "")
            try:
                d[k] = float(v) * (scale(v, units[i]) if k in ("dram_read", "dram_write", "time_us") else 1)
            except ValueError:
                d[k] = v
    return d


out_prefix = sys.argv[1]
forms = {}
for arg in sys.argv[2:]:
    name, rest = arg.split("=")
    path, pairs, ar, aw = rest.split(":")
    d = load(path)
    d.update(pairs=int(pairs), algorithmic_read=int(ar), algorithmic_write=int(aw))
    forms[name] = d
json.dump({"source": "ncu --set full --import-source on --clock-control none, one launch per form (cold-cache replay): scripts/r2_measure.sh", "forms": forms},
          open(out_prefix + ".json", "w"), indent=1)
with open(out_prefix + ".md", "w") as f:
    f.write("| form | kernel | units | grid x block | regs | time us (ncu, cold) | DRAM read MB | DRAM write MB | algorithmic read / write MB | warps active % | sm throughput % |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
    for n, d in forms.items():
        f.write("| %s | `%s` | %d | %s x %s | %s | %.1f | %.2f | %.2f | %.2f / %.2f | %.1f | %.1f |\n" % (
            n, d["kernel"], d["pairs"], d["grid"], d["block"], int(d["registers"]), d["time_us"], d["dram_read"] / 1e6, d["dram_write"] / 1e6,
            d["algorithmic_read"] / 1e6, d["algorithmic_write"] / 1e6, d["warps_active_pct"], d["sm_throughput_pct"]))
print(open(out_prefix + ".md").read())
