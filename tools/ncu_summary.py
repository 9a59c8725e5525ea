#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report as markdown:  ncu_summary.py <report.ncu-rep> [title]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def main(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n\nsource: `{path}` (`ncu --set full --clock-control none`)\n")
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        name = d.get("Kernel Name", ("", "?"))[1]
        print(f"## `{name[:110]}`\n\n| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            for h in hdr:
                if h == k or h.endswith("." + k):
                    u, v = d[h]
                    if v != "":
                        print(f"| {k} | {v} | {u} |")
                    break
        try:   # tcgen05 issue time: the hmma sub-pipe counter sums the SM's 4 sub-partitions; the pct metric above does not track it
            hm = float(d[[h for h in hdr if h.endswith("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg")][0]][1].replace(",", ""))
            cy = float(d[[h for h in hdr if h.endswith("sm__cycles_elapsed.max")][0]][1].replace(",", ""))
            print(f"| derived: hmma sub-pipe active / (4 x elapsed cycles) | {hm / (4 * cy) * 100:.1f} | % |")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu summary")
