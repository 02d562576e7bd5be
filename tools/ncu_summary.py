"""Condense an ncu report (--page raw --csv) into the small CSV kept under profiles/."""
import csv
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second', 'smsp__pcsamp_sample_count']


def main(rep, out_csv, traffic_json=None, label=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    stall = sorted(k for k in d if k.startswith('smsp__pcsamp_warps_issue_stalled') and not k.endswith('_not_issued'))
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        w.writerow(["kernel", d.get("Kernel Name", ("", ""))[0], label])
        for k in KEYS + stall:
            if k in d:
                w.writerow([k, d[k][0], d[k][1]])
    if traffic_json:
        def tobytes(key):
            v, u = d[key]
            return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        rd, wr = tobytes('dram__bytes_read.sum'), tobytes('dram__bytes_write.sum')
        def num(key):
            return float(d[key][0]) if key in d else None
        json.dump({"kernel": d.get("Kernel Name", ("", ""))[0] + " " + label, "dram_bytes_per_launch": rd + wr,
                   "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "lsu_shared_wavefronts_per_launch": (num('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum') or 0)
                   + (num('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum') or 0),
                   "fma_pipe_pct": num('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed'),
                   "issue_active_pct": num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                   "shared_wavefront_pct": num('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
                   "duration_us": num('gpu__time_duration.sum'),
                   "source": f"ncu --set full --clock-control none ({rep})"}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
