"""Condenses an `ncu --set full` report of one cb200_ecm_device call (tools/ecm_once.py) into the JSON kept
under profiles/ and read by bench.py for `roofline.traffic` (run where ncu is installed; launches nothing).

usage: python tools/ncu_summary2.py OUT.json M N "how the reports were captured" REPORT.ncu-rep [REPORT.ncu-rep ...]

Per kernel family (the names bench.py / cb200_ctx_kernel_ms use): launches seen, mean duration, DRAM bytes per
launch, registers, occupancy, pipe / issue utilisation, stall reasons per issued instruction, and the
algorithmic bytes per launch at the captured shape (bench.algorithmic_bytes)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

FAMILY = [("fold_kernel", "fold"), ("lean_fwd_compose", "forward_compose"), ("lean_fwd_replay", "forward_scan"),
          ("lean_bwd_replay_kernel<1, 0>", "backward_scan"), ("lean_bwd_replay_kernel<0, 1>", "backward_publish"),
          ("lean_bwd_replay_kernel<(bool)1, (bool)0>", "backward_scan"),
          ("lean_bwd_replay_kernel<(bool)0, (bool)1>", "backward_publish"),
          ("lean_group_scan", "segment_scan"), ("residual_kernel", "residuals"), ("lean_gather", "precision_updates"),
          ("lean_scatter", "precision_updates")]
RAW = {
    "time_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "registers": "launch__registers_per_thread",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "warp_instructions": "smsp__inst_executed.sum",
}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def main():
    out_path, m, n, how, reports = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5:]
    alg = bench.algorithmic_bytes(m, [n], True, lambda _n: 5)
    fams = {}
    for report in reports:
        out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, rows = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        accumulate(fams, hdr, units, rows, col)
    finish(fams, alg, m, n, how, out_path)


def accumulate(fams, hdr, units, rows, col):
    for r in rows:
        name = r[col["Kernel Name"]]
        fam = next((f for tag, f in FAMILY if tag in name), None)
        if fam is None:
            continue
        k = fams.setdefault(fam, {"kernels": set(), "n": 0, "acc": {}, "stall": {}})
        k["kernels"].add(name)
        k["n"] += 1
        for short, metric in RAW.items():
            if metric in col and r[col[metric]] != "":
                v = float(r[col[metric]].replace(",", "")) * UNIT.get(units[col[metric]], 1.0)
                k["acc"][short] = k["acc"].get(short, 0.0) + v
        for i, h in enumerate(hdr):
            mt = STALL.match(h)
            if mt and r[i] not in ("", "0"):
                k["stall"][mt.group(1)] = k["stall"].get(mt.group(1), 0.0) + float(r[i])


def finish(fams, alg, m, n, how, out_path):
    kernels = {}
    for fam, k in fams.items():
        e = {short: v / k["n"] for short, v in k["acc"].items()}
        e["kernels"] = sorted(k["kernels"])
        e["launches_captured"] = k["n"]
        e["dram_bytes_per_launch"] = e.get("dram_read_bytes", 0.0) + e.get("dram_write_bytes", 0.0)
        e["stall_per_issue"] = {s: round(v / k["n"], 3) for s, v in sorted(k["stall"].items(), key=lambda kv: -kv[1])[:8]}
        if fam in alg and alg[fam][1] > 0:
            e["algorithmic_bytes_per_launch"] = alg[fam][0] / alg[fam][1]
            e["dram_over_algorithmic"] = e["dram_bytes_per_launch"] / e["algorithmic_bytes_per_launch"]
            e["thread_instructions_per_interval"] = e.get("warp_instructions", 0.0) * 32 / n
        e["shape"] = f"{m} tracks x {n} intervals"
        kernels[fam] = e
    json.dump({"source": how, "kernels": kernels}, open(out_path, "w"), indent=1)
    print(json.dumps({f: {x: round(v[x], 3) for x in ("time_us", "dram_bytes_per_launch", "registers", "issue_active_pct",
                                                     "fp64_pipe_pct", "dram_over_algorithmic") if x in v}
                      for f, v in kernels.items()}, indent=0))


if __name__ == "__main__":
    main()
