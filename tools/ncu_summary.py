#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports: one line of the roofline-relevant metrics per captured launch."""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]
for path in sys.argv[1:]:
    with open(path) as f:
        rows = list(csv.reader(f))
    if len(rows) < 3:
        print(path, "empty"); continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print(f"== {path}: {d.get('Kernel Name','?')[:80]}")
        for k in KEYS:
            if k in d: print(f"   {k:62s} {d[k]:>16s} {u[k]}")
        try:
            t = float(d["gpu__time_duration.sum"].replace(",", "")); ut = u["gpu__time_duration.sum"]
            t *= {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1, "second": 1, "nsecond": 1e-9}[ut]
            def b(k):
                v = float(d[k].replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
            tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
            print(f"   traffic = {tr/1e9:.4f} GB  in {t*1e3:.3f} ms  =>  {tr/t/1e9:.0f} GB/s DRAM")
        except Exception as e:
            print("   (traffic n/a:", e, ")")
