"""Summaries of an ncu report for profiles/: one CSV row per captured launch with the metrics the README tables quote,
and (optionally) the DRAM bytes per launch that bench.py reports as roofline.traffic.

    python tools/ncu_extract.py gpurun_out/r02_conv_step.ncu-rep profiles/r02_conv_step_metrics.csv [--traffic c2]
"""
import csv
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(m) for m in METRICS if m in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name"] + [hdr[c] for c in cols])
        w.writerow(["", ""] + [units[c] for c in cols])
        for r in rows[2:]:
            w.writerow([r[hdr.index("ID")], r[hdr.index("Kernel Name")][:120]] + [r[c] for c in cols])
    if "--traffic" in sys.argv:
        cfg = sys.argv[sys.argv.index("--traffic") + 1]
        # launch order of one step's tensor-core kernels (tools/prof_r02.sh: -k conv_tc|wgrad_tc|conv0_win_fwd -c 7)
        order = ["conv0.fwd_fused", "conv1.fwd", "conv2.fwd", "conv2.wgrad", "conv2.dgrad", "conv1.wgrad", "conv1.dgrad"]
        rd, wr, nm = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")

        def to_bytes(v, u):
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        launches = rows[2:2 + len(order)]
        # wgrad / dgrad order inside a block depends on the stream layout of the build: name the phases by kernel
        tr, seen = {}, {"conv": 0, "wgrad": 0}
        for r in launches:
            n = r[nm]
            b = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
            if "conv0_win_fwd" in n:
                tr["conv0.fwd_fused"] = b
            elif "wgrad_tc" in n:
                tr[["conv2.wgrad", "conv1.wgrad"][seen["wgrad"]]] = b
                seen["wgrad"] += 1
            elif "conv_tc" in n:
                tr[["conv1.fwd", "conv2.fwd", "conv2.dgrad", "conv1.dgrad"][seen["conv"]]] = b
                seen["conv"] += 1
        path = os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_traffic.json")
        try:
            with open(path) as f:
                allcfg = json.load(f)
        except Exception:
            allcfg = {}
        allcfg[cfg] = tr
        allcfg["_source"] = (f"profiles/{os.path.basename(out)} (ncu --set full of one step's conv_tc_kernel / wgrad_tc_kernel / "
                             "conv0_win_fwd_kernel launches, tools/prof_r02.sh + tools/ncu_extract.py): dram__bytes_read.sum + "
                             "dram__bytes_write.sum per launch.  Writes below the algorithmic output size are lines still "
                             "dirty in the 126 MB L2 when the kernel ends.")
        with open(path, "w") as f:
            json.dump(allcfg, f, indent=1)


if __name__ == "__main__":
    main()
