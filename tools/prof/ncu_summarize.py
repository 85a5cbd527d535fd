"""Summarise an .ncu-rep (ncu --set full capture) into the handful of metrics profiles/*.md quote.
  python tools/prof/ncu_summarize.py gpurun_out/x.ncu-rep > profiles/x_summary.csv"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow(["launch", "kernel", "metric", "value", "unit"])
    for li, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].split("(")[0]
        for m in WANT:
            if m in idx:
                w.writerow([li, name, m, r[idx[m]], units[idx[m]]])


if __name__ == "__main__":
    main(sys.argv[1])
