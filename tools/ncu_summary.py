"""Summarises an .ncu-rep (read here, without a GPU: `ncu -i … --page raw --csv`) into a small JSON:
per captured launch the kernel name, duration, DRAM bytes, issue-slot utilisation, executed warp
instructions, occupancy, registers and the leading warp-stall reasons.

  python tools/ncu_summary.py gpurun_out/r2_prof_integrate.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}
UNIT_SCALE = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        d = {"kernel": r[col["Kernel Name"]][:110]}
        for k, name in KEYS.items():
            if k in col and r[col[k]] != "":
                v = float(r[col[k]].replace(",", ""))
                u = units[col[k]]
                if u in UNIT_SCALE and (name.endswith("_us") or name.endswith("_MB")):
                    v *= UNIT_SCALE[u]
                d[name] = round(v, 3)
        # warps stalled per issue slot, by reason (ratio per issue-active cycle)
        stalls = {h.split("issue_stalled_")[1].split("_per_issue_active")[0]: float(r[i].replace(",", ""))
                  for h, i in col.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] != ""}
        d["top_stalls_warps_per_issue"] = {k: round(v, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]}
        out.append(d)
    js = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(js)
    print(js)


if __name__ == "__main__":
    main()
