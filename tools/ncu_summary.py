"""ncu raw CSV pages (ncu -i X.ncu-rep --page raw --csv) -> a short table and profiles/r02_traffic.json (what bench.py's
`roofline.traffic` reads: dram__bytes_read.sum + dram__bytes_write.sum per launch of each kernel of the encode step).
    python tools/ncu_summary.py profiles/r02_prof_encode.raw.csv [n_sent] [--write-traffic]"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def rows(path):
    r = list(csv.reader(open(path)))
    hdr, units = r[0], r[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for row in r[2:]:
        name = re.sub(r"<unnamed>::|\(.*", "", row[idx["Kernel Name"]]).strip()
        out = {"kernel": name}
        for w in WANT:
            if w in idx:
                v = float(row[idx[w]].replace(",", ""))
                out[w] = v * UNIT.get(units[idx[w]], 1)
        yield out


def main():
    path = sys.argv[1]
    n_sent = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 1_000_000
    kernels = {}
    for r in rows(path):
        dram = r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)
        t = r.get("gpu__time_duration.sum", 0)
        print("%-28s %8.1f us  dram %7.1f MB (%5.0f GB/s)  issue %4.1f %%  L2 hit %4.1f %%  regs %d  warp-inst %.0f M" % (
            r["kernel"], t * 1e6, dram / 1e6, dram / t / 1e9 if t else 0, r.get("sm__inst_issued.avg.pct_of_peak_sustained_active", 0),
            r.get("lts__t_sector_hit_rate.pct", 0), r.get("launch__registers_per_thread", 0), r.get("smsp__inst_executed.sum", 0) / 1e6))
        kernels[r["kernel"]] = {"dram_bytes": int(dram), "duration_us": round(t * 1e6, 1)}
    if "--write-traffic" in sys.argv:
        out = {"n_sent": n_sent, "source": os.path.relpath(path, ROOT) + " (ncu --set full, one launch each, same workload as the bench line)",
               "kernels": kernels}
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as f:
            json.dump(out, f, indent=1)
        print("wrote profiles/r02_traffic.json")


if __name__ == "__main__":
    main()
