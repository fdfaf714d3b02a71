"""Per-launch summary of an `ncu --set full` capture, plus the DRAM traffic of the dominant kernel.

    python profiles/tools/ncu_summarise.py gpurun_out/prof_r2_wave.ncu-rep profiles/r02_wave_ncu_full_summary.csv \
        [--skip N]   (N leading launches are the warm-up call of profiles/tools/ncu_wave.py)
        [--no-traffic]   (a capture of something else than the predict wave, e.g. ncu_train.py:
                          leave profiles/roofline_traffic.json alone)

Writes the CSV (one row per kernel launch: duration, DRAM bytes, tensor-pipe / shared-memory /
L2 / SM utilisation, registers, dynamic shared memory, cluster size, achieved occupancy) and
rewrites profiles/roofline_traffic.json (mean dram read+write bytes per launch of
conv3x3_zfold2_kernel, the kernel bench.py's `roofline` describes) from the same capture."""

import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

METRICS = [
    "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "smsp__inst_executed.sum",
    "nvlink__bytes_tx.sum", "nvlink__bytes_rx.sum",
]


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, body = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    keep = ["ID", "Kernel Name"] + [m for m in METRICS if m in col]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[col[k]] for k in keep])
        for r in body[skip:]:
            w.writerow([r[col[k]] for k in keep])
    # DRAM traffic of the dominant kernel, per launch
    def num(r, k):
        v = float(r[col[k]].replace(",", "")) if r[col[k]] else 0.0
        u = units[col[k]].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    per = [num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum") for r in body[skip:]
           if "conv3x3_zfold2_kernel" in r[col["Kernel Name"]]]
    if per and "--no-traffic" not in sys.argv:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as f:
            json.dump({"kernel": "conv3x3_zfold2_kernel",
                       "source": f"{os.path.relpath(out_csv, ROOT)} (ncu --set full --clock-control none, "
                                 f"{len(per)} launches = one wave of 32 patches)",
                       "traffic_bytes_per_launch": sum(per) / len(per), "per_launch_bytes": per}, f, indent=1)
    print(f"{len(body) - skip} launches -> {out_csv}; zfold2 launches: {len(per)}")


if __name__ == "__main__":
    main()
