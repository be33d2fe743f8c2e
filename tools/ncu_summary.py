#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the JSON kept under profiles/ (per-launch numbers of every captured kernel).

  python tools/ncu_summary.py REPORT.ncu-rep OUT.json --windows N --command "<the ncu command line>" [--sweep]

With --sweep the DRAM bytes of the Jacobian-mode sweep kernels (k_proj<1,..>, k_line_vp<1>, k_imu_geom<1>, k_imu_weight,
the first k_prior) are added up into `jacobian_sweep_dram_bytes` (what bench.py reports as roofline.traffic).
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "time_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "registers": "launch__registers_per_thread",
    "fp64_pipe_pct": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum",
    "tensor_inst": "sm__inst_executed_pipe_tensor.sum",
    "tensor_pipe_dmma_pct": "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    windows = int(sys.argv[sys.argv.index("--windows") + 1]) if "--windows" in sys.argv else None
    command = sys.argv[sys.argv.index("--command") + 1] if "--command" in sys.argv else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels, sweep = {}, 0.0
    seen_prior = False
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("uvs::", "")
        if "<" in name:
            short = name[: name.index(">") + 1].replace("void ", "").replace("uvs::", "").replace("(bool)", "")
        ent = {"grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for k, m in KEYS.items():
            if m in idx and r[idx[m]] not in ("", "n/a"):
                v = float(r[idx[m]].replace(",", ""))
                ent[k] = v * UNIT_SCALE.get(units[idx[m]], 1.0)
        key = short
        n = 2
        while key in kernels:
            key = "%s #%d" % (short, n)
            n += 1
        kernels[key] = ent
        if "--sweep" in sys.argv:
            jac = ("k_proj<1" in short or "k_line_vp<1" in short or "k_imu_geom<1" in short or short == "k_imu_weight"
                   or (short == "k_prior" and not seen_prior))
            if short == "k_prior":
                seen_prior = True
            if jac and key == short:
                sweep += ent.get("dram_read_bytes", 0.0) + ent.get("dram_write_bytes", 0.0)
    doc = {"command": command, "windows": windows,
           "note": "per launch; dram write bytes under-count what a kernel produces because dirty lines still sit in the 126 MB L2 when it ends",
           "kernels": kernels}
    if "--sweep" in sys.argv:
        import os
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        doc["jacobian_sweep_dram_bytes"] = sweep
        doc["kernel_sources_sha"] = bench.sources_sha()   # bench.py refuses the file when the sweep sources have changed since
    json.dump(doc, open(out, "w"), indent=1)
    print("wrote %s: %d kernels" % (out, len(kernels)))


if __name__ == "__main__":
    main()
