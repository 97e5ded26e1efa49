"""Summarise ncu outputs into profiles/ (run in the build container, reads gpurun_out/).

  python tools/ncu_summarize.py launches gpurun_out/r01_launches.csv profiles/r01_tc_launches_summary.csv "<command>"
  python tools/ncu_summarize.py full gpurun_out/r01_tc_full.ncu-rep profiles/r01_tc_ncu_full_summary.json <units> "<command>" [config kernel_id samples batch]
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "sm__cycles_active.max", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * mult.get(unit, 1)


def launches(src, dst, command):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v if r[iu] in ("ms", "msecond") else v * 1e3
        name = r[ik].split("(")[0][:70]
        tot[name][0] += 1
        tot[name][1] += ms
    total = sum(t[1] for t in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (command: {command})\n# per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{ms:.3f},{ms / total:.4f}\n")
    print(open(dst).read())


def full(src, dst, units, command, config=None, kernel_id=None, samples=None, batch=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, unit_row, val = rows[0], rows[1], rows[2]
    m = {}
    for name in KEEP:
        if name in hdr:
            i = hdr.index(name)
            m[name] = {"value": val[i], "unit": unit_row[i]}
    stalls = {}
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(val[i])
            except ValueError:
                pass
    top = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    rd = to_bytes(m["dram__bytes_read.sum"]["value"], m["dram__bytes_read.sum"]["unit"])
    wr = to_bytes(m["dram__bytes_write.sum"]["value"], m["dram__bytes_write.sum"]["unit"])
    res = {"kernel": val[hdr.index("Kernel Name")], "capture": command, "units_in_capture": units,
           "config": config, "kernel_id": kernel_id, "capture_samples": samples, "capture_batch": batch,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_unit": round((rd + wr) / units, 2),
           "metrics": {**m, "top_stalls_per_issue": top}}
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "metrics"}), json.dumps(m)[:1500], top)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        extra = sys.argv[6:10]
        full(sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5], *( [extra[0], int(extra[1]), int(extra[2]), int(extra[3])] if len(extra) == 4 else []))
