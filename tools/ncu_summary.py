"""Text summary of one `ncu --set full` capture of the step kernel for profiles/: headline metrics and stall reasons
from the raw page, by-function / by-line table from tools/ncu_lines.py.
usage: ncu_summary.py raw.csv lines.txt "<header line>" "<command line>" n_envs > profiles/<name>.txt (+ traffic json on stderr)"""
import csv, json, sys
raw, lines, header, command, n_envs = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
rows = list(csv.reader(open(raw)))
h, units, v = rows[0], rows[1], rows[2]
g = lambda k: (v[h.index(k)], units[h.index(k)]) if k in h else ("n/a", "")
print("# " + header)
print("# command: " + command)
for k in ("gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__inst_executed.avg.per_cycle_active",
          "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
          "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.per_cycle_active",
          "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"):
    print("%-70s %s %s" % (k, *g(k)))
print("\n# warp stall reasons (cycles per issued instruction)")
st = []
for i, n in enumerate(h):
    if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio") and "not_issued" not in n:
        try:
            st.append((float(v[i]), n))
        except ValueError:
            pass
for val, n in sorted(st, reverse=True)[:10]:
    print("%-90s %.3f" % (n, val))
print()
print(open(lines).read())
def num(k):
    val, unit = g(k)
    return float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(unit, 1)
rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
json.dump({"source": "ncu --set full, %d envs" % n_envs, "envs": n_envs, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "bytes_per_env_step": (rd + wr) / n_envs}, sys.stderr)
