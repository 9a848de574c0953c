"""Summarise an `ncu --page raw --csv` export (one kernel launch): stalls per issue, pipes, DRAM bytes."""
import csv
import sys


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    vals = rows[-1]
    d = dict(zip(hdr, vals))

    def f(k):
        try:
            return float(d[k].replace(",", ""))
        except Exception:
            return None
    out = []
    out.append(f"kernel: {d.get('Kernel Name')}  block {d.get('Block Size')} grid {d.get('Grid Size')}")
    for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
              "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
              "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
              "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed"):
        if k in d:
            out.append(f"  {k} = {d[k]}")
    st = {}
    for k, v in d.items():
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
            try:
                st[k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(v.replace(",", ""))
            except ValueError:
                pass
    out.append("  stalls per issue: " + ", ".join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(summarise(p))
