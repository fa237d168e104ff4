"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON for profiles/.
    python tools/ncu_summary.py gpurun_out/x/prof.ncu-rep profiles/name.json"""
import csv
import io
import json
import subprocess
import sys

WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
smsp__thread_inst_executed_per_inst_executed.ratio smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__warps_eligible.avg.per_cycle_active sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
l1tex__t_sector_hit_rate.pct l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct lts__t_sector_hit_rate.pct
sm__cycles_elapsed.avg lts__t_bytes.sum""".split()


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for i, h in enumerate(hdr):
            if h in ("Kernel Name", "Block Size", "Grid Size"):
                d[h] = vals[i]
            if h in WANT or (("issue_stalled" in h) and h.endswith("per_warp_active.pct")):
                d[h] = {"unit": units[i], "value": vals[i]}
        res.append(d)
    json.dump({"report": rep, "kernels": res}, open(out, "w"), indent=1)
    for d in res:
        print(d.get("Kernel Name"))
        for k, v in d.items():
            if isinstance(v, dict):
                print("  %-80s %s %s" % (k, v["value"], v["unit"]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
