"""Condenses an `ncu --set full --import-source on` report of the scan kernels into the JSON kept
under profiles/ (run where ncu is installed; reads the report, launches nothing).

usage: python tools/ncu_summary.py REPORT.ncu-rep OUT.json ["how the report was captured"]

Per kernel: duration, DRAM bytes, registers, pipe / issue utilisation, stall reasons per issued
instruction, and -- from the source page joined with the line table of the shipped library -- the
thread instructions executed per interval, attributed to the source function they were inlined from.
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "consenrich_b200", "lib", "libconsenrich_b200.so")
CSRC = os.path.join(ROOT, "consenrich_b200", "csrc")
N_BINS = 2344705  # bench.py's workload: instructions are reported per interval of it

RAW = {
    "time_us": "gpu__time_duration.sum",
    "dram_read_MB": "dram__bytes_read.sum",
    "dram_write_MB": "dram__bytes_write.sum",
    "grid": "launch__grid_size",
    "registers": "launch__registers_per_thread",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_read_pct_of_peak": "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "dram_write_pct_of_peak": "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
}
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def ncu_csv(report, page, extra=()):
    out = subprocess.run(["ncu", "-i", report, "--page", page, "--csv", *extra], check=True, capture_output=True,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def line_table(kernel_tag):
    """address -> (opcode, (file, line)) of one kernel of the shipped library (nvdisasm -g)."""
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.startswith("ssm_kernels") and "cabi" not in f][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], check=True, capture_output=True,
                             text=True).stdout
    name = [l for l in re.findall(r"\.text\.(\S+?)[:,]", txt) if kernel_tag in l][0]
    i = txt.index(".text." + name + ":")
    j = txt.find("//--------------------- .text.", i + 10)
    cur, table = None, {}
    for line in txt[i:j if j > 0 else None].split("\n"):
        m = re.match(r'\s*//## File "(.*?)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            table[int(m.group(1), 16)] = (m.group(3), cur)
    return table


def function_starts(path):
    out = []
    for n, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"\s*(?:template.*>\s*)?(?:CB_HD_COLD|CB_HD|__device__|__global__|static|inline|__forceinline__|\s)+"
                     r"[\w:<>\*& ]+?\b(\w+)\s*\(", l)
        if m and not l.strip().startswith(("return", "if", "for", "while", "//", "const ", "double ", "int ", "else")):
            out.append((n, m.group(1)))
    return out


FUNCS = {f: function_starts(os.path.join(CSRC, f)) for f in ("ssm_math.cuh", "ssm_kernels.cu")}


def function_of(loc):
    if loc is None:
        return "?"
    f, l = loc
    if f not in FUNCS:
        return f
    name = "?"
    for n, nm in FUNCS[f]:
        if n > l:
            break
        name = nm
    return name


def main():
    report, out_path = sys.argv[1], sys.argv[2]
    how = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = ncu_csv(report, "raw")
    hdr, rows = raw[0], raw[2:]
    col = {h: i for i, h in enumerate(hdr)}
    kernels = {}
    for r in rows:
        kname = r[col["Kernel Name"]]
        key = "forward_scan" if "Fwd" in kname else "backward_scan" if "Bwd" in kname else kname
        k = {"kernel": kname}
        for short, metric in RAW.items():
            if metric in col and r[col[metric]] != "":
                k[short] = float(r[col[metric]].replace(",", ""))
        k["dram_bytes_per_launch"] = (k.get("dram_read_MB", 0.0) + k.get("dram_write_MB", 0.0)) * 1e6
        k["stall_per_issue"] = {m.group(1): round(float(r[i]), 3) for i, h in enumerate(hdr) for m in [STALL.match(h)]
                                if m and r[i] not in ("", "0")}
        kernels[key] = k
    # source page: executed instructions per function
    src = ncu_csv(report, "source", ("--print-source", "sass"))
    sections, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            sections.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r and re.match(r"^(0x)?[0-9a-f]+$", r[0] or "z"):
            cur["rows"].append(r)
    seen = set()
    for sec in sections:
        key = "forward_scan" if "Fwd" in sec["name"] else "backward_scan" if "Bwd" in sec["name"] else sec["name"]
        if key in seen or key not in kernels:
            continue
        seen.add(key)
        tag = re.sub(r"[^A-Za-z0-9]", "", "Fwd2ILb1EEELb0" if "Fwd2<(bool)1>" in sec["name"] else
                     "Bwd2ILb1EEELb0" if "Bwd2<(bool)1>" in sec["name"] else "")
        if not tag:
            continue
        table = line_table(tag)
        h = sec["hdr"]
        ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
        addr = lambda r: int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        base = min(addr(r) for r in sec["rows"])
        inst, samp = collections.Counter(), collections.Counter()
        for r in sec["rows"]:
            f = function_of(table.get(addr(r) - base, ("?", None))[1])
            inst[f] += int(r[ie] or 0)
            samp[f] += int(r[isamp] or 0)
        tot, tots = sum(inst.values()), max(sum(samp.values()), 1)
        kernels[key]["thread_instructions_per_interval"] = round(tot * 32 / N_BINS, 1)
        kernels[key]["by_function"] = [
            {"function": f, "thread_instructions_per_interval": round(v * 32 / N_BINS, 1),
             "share_of_instructions_pct": round(100 * v / tot, 1), "share_of_stall_samples_pct": round(100 * samp[f] / tots, 1)}
            for f, v in inst.most_common(16)]
    json.dump({"source": how, "kernels": kernels}, open(out_path, "w"), indent=1)
    print(json.dumps({k: {x: v[x] for x in ("time_us", "dram_bytes_per_launch", "registers") if x in v}
                      for k, v in kernels.items()}))


if __name__ == "__main__":
    main()
