#!/usr/bin/env python
"""Turn an Nsight Compute report (.ncu-rep) into the small JSON summary kept under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.json"""
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        e = {"Kernel Name": d.get("Kernel Name", "").replace("gimb::<", "").replace("void ", "void ")}
        for k in KEYS:
            if k in d:
                e[k] = (d[k] + " " + u.get(k, "")).strip()
        res.append(e)
    json.dump(res, open(out, "w"), indent=1)
    print(f"{len(res)} kernels -> {out}")


if __name__ == "__main__":
    main()
