"""Condense `ncu --set full --csv --page raw` captures (one kernel each) into metric,unit,value tables for profiles/.

  python tools/condense_ncu.py gpurun_out/r2_full_*.csv   ->  profiles/r2_ncu_full_<name>.csv   (prefix: $NCU_PREFIX, default r2)
"""
import os
import csv
import re
import sys
from pathlib import Path

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|dram__throughput\.avg\.pct|"
    r"gpu__dram_throughput\.avg\.pct|lts__t_bytes\.sum(\.per_second)?|lts__t_sector_hit_rate\.pct|"
    r"l1tex__t_bytes\.sum(\.per_second)?|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__inst_executed_pipe_(tensor|xu|fma|alu|lsu|uniform|tma|tc)[a-z_0-9]*\.(sum|avg\.pct_of_peak_sustained_active)|"
    r"sm__pipe_(tensor|xu|fma|alu)[a-z_0-9]*cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|smsp__cycles_active\.avg|"
    r"smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active\.ratio|smsp__issue_active\.avg\.pct|"
    r"launch__(grid_size|block_size|cluster_size|cluster_max_active|registers_per_thread|shared_mem_per_block|"
    r"occupancy_limit_[a-z_]+|waves_per_multiprocessor)|sm__maximum_warps_per_active_cycle_pct|"
    r"smsp__sass_inst_executed_op_(shared|global|local)_[a-z]+\.sum|sass__inst_executed_local_(loads|stores))")

out_dir = Path(__file__).resolve().parent.parent / "profiles"
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    kname = vals[names.index("Kernel Name")]
    tag = Path(f).stem.replace("r2_full_", "").replace("full_", "")
    with open(out_dir / f"{os.environ.get('NCU_PREFIX', 'r2')}_ncu_full_{tag}.csv", "w") as o:
        o.write("metric,unit,value\n")
        o.write(f"Kernel Name,,{kname.replace(',', ';')}\n")
        for n, u, v in zip(names, units, vals):
            if KEEP.match(n):
                o.write(f"{n},{u},{v.replace(',', '')}\n")
    d = dict(zip(names, vals))
    rd, wr = float(d["dram__bytes_read.sum"].replace(",", "")), float(d["dram__bytes_write.sum"].replace(",", ""))
    ur, uw = units[names.index("dram__bytes_read.sum")], units[names.index("dram__bytes_write.sum")]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    print(tag, kname[:70], "time", d["gpu__time_duration.sum"], units[names.index("gpu__time_duration.sum")],
          "dram bytes", int(rd * scale[ur] + wr * scale[uw]))
