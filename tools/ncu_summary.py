"""Summarise .ncu-rep captures (read with `ncu -i ... --page raw --csv`) into a small markdown table.

    python tools/ncu_summary.py gpurun_out/prof_apply_tma.ncu-rep ... > profiles/r01_ncu_summary.md
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes.sum.per_second", "DRAM throughput"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy %"),
    ("sm__inst_executed_pipe_tma.sum", "TMA instructions"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "LSU global-load sectors"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    for vals in r[2:]:
        yield {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def main():
    print("| capture | kernel | " + " | ".join(label for _, label in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for path in sys.argv[1:]:
        for row in rows_of(path):
            name = row.get("Kernel Name", ("?", ""))[0].split("(")[0].replace("void ", "")
            cells = []
            for key, _ in KEYS:
                v, u = row.get(key, ("", ""))
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            print(f"| {path.split('/')[-1]} | `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
