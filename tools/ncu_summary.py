"""Condense an `ncu --set full` report into the JSON kept under profiles/ and refresh
profiles/ncu_traffic_latest.json (read by bench.py for roofline.traffic / issue_roofline).

    python tools/ncu_summary.py gpurun_out/prof_render.ncu-rep profiles/r01_ncu_full_render_kernels_v4.json

Runs on the CPU box: `ncu -i <rep> --page raw --csv` needs no GPU.
"""
from __future__ import annotations

import csv
import io
import json
import subprocess
import sys
from pathlib import Path

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_shared_atom.sum", "sm__inst_executed_pipe_xu.sum", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
]
_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(header)}
    summary, traffic = [], {}
    for r in body:
        name = r[col["Kernel Name"]]
        entry = {"Kernel Name": name}
        for k in KEEP:
            if k in col:
                entry[k] = f"{r[col[k]]} {units[col[k]]}".strip()
        summary.append(entry)
        short = name.split("(")[0].split("<")[0].replace("void ", "").strip()  # "void k_render_bwd<1>(...)" -> k_render_bwd

        def to_bytes(key):
            return float(r[col[key]].replace(",", "")) * _BYTES.get(units[col[key]], 1.0)

        traffic[short] = {
            "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
            "warp_instructions_per_launch": float(r[col["smsp__inst_executed.sum"]].replace(",", "")),
            "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]].replace(",", "")),
            "source": f"{out} (ncu --set full --clock-control none, tools/profile_step.py)",
        }
    Path(out).write_text(json.dumps(summary, indent=1))
    latest = Path(__file__).resolve().parents[1] / "profiles" / "ncu_traffic_latest.json"
    merged = json.loads(latest.read_text()) if latest.exists() else {}
    merged.update(traffic)  # a report usually holds a subset of the kernels: keep the others' latest entries
    latest.write_text(json.dumps(merged, indent=1))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
