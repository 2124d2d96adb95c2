#!/usr/bin/env python
"""Reduce `ncu -i capture.ncu-rep --page raw --csv` (a `--set full` capture: ~2000 columns) to the columns the
records under profiles/ keep -- the format `bench.py::ncu_traffic_per_launch` reads.

  ncu -i gpurun_out/x.ncu-rep --page raw --csv | python tools/ncu_compact.py > profiles/rNN_ncu_full_x.csv
"""
import csv
import sys

KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__block_size", "launch__cluster_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum"]


def compact(rows):
    rows = [r for r in rows if len(r) > 10]
    head, units, body = rows[0], rows[1], rows[2:]
    name = head.index("Kernel Name")
    cols = [(k, head.index(k)) for k in KEEP if k in head]
    out = [["Kernel Name"] + [k for k, _ in cols], [""] + [units[i] for _, i in cols]]
    for r in body:
        out.append([r[name]] + [r[i].replace(",", "") for _, i in cols])
    return out


if __name__ == "__main__":
    csv.writer(sys.stdout, lineterminator="\n").writerows(compact(list(csv.reader(sys.stdin))))
