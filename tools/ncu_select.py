"""Selected counters of an `ncu --page raw --csv` export -> a small JSON for profiles/ (the full .ncu-rep stays in gpurun_out/).
   python tools/ncu_select.py gpurun_out/X_raw.csv profiles/X_raw_selected.json [row]"""
import csv, json, re, sys
KEEP = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|waves_per_multiprocessor|shared_mem_per_block_dynamic)|"
                  r"sm__(inst_executed_pipe_(xu|fma|alu|fp64|lsu|uniform|tensor\w*)\.avg\.pct_of_peak_sustained_active|pipe_(tensor|fma|fp64|alu)_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"warps_active\.avg\.pct_of_peak_sustained_active|throughput\.avg\.pct_of_peak_sustained_elapsed|issue_active\.avg\.pct_of_peak_sustained_elapsed|cycles_active\.avg)|"
                  r"smsp__(issue_active\.avg\.(pct_of_peak_sustained_active|per_cycle_active)|warps_eligible\.avg\.per_cycle_active|warps_active\.avg\.per_cycle_active|inst_executed\.sum|"
                  r"average_warps_issue_stalled_\w+_per_issue_active\.ratio)|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared(_op_(ld|st))?\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared(_op_(ld|st))?\.sum|lts__t_sector_hit_rate\.pct|"
                  r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|sm__sass_inst_executed_op_shared_(ld|st)\.sum)$")
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
row = rows[hdr + 2 + (int(sys.argv[3]) if len(sys.argv) > 3 else 0)]
out = {"kernel": row[names.index("Kernel Name")], "source": sys.argv[1]}
for i, n in enumerate(names):
    if KEEP.match(n):
        v = row[i].replace(",", "")
        try:
            v = float(v)
        except ValueError:
            pass
        out[n] = {"value": v, "unit": units[i]}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print("wrote", sys.argv[2], len(out) - 2, "metrics; kernel", out["kernel"][:90])
