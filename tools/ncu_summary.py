"""Summarises an ncu report (`ncu --set full`) into JSON: one record per profiled launch with the counters the
design decisions cite. Usage: python tools/ncu_summary.py report.ncu-rep [out.json]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_sector_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_sector_hit_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct_of_peak",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__occupancy_limit_shared_mem": "blocks_per_sm_limit_smem",
    "launch__occupancy_limit_registers": "blocks_per_sm_limit_regs",
    "launch__grid_size": "grid",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "global_load_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "global_load_requests",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        rec = {"kernel": r[ix["Kernel Name"]]}
        for k, name in WANT.items():
            if k in ix and r[ix[k]] != "":
                v = float(r[ix[k]].replace(",", ""))
                u = units[ix[k]]
                rec[name] = v * SCALE[u] if u in SCALE else v
        if "dram_read" in rec:
            rec["dram_bytes_per_launch"] = rec["dram_read"] + rec.get("dram_write", 0.0)
            rec["dram_gbs"] = rec["dram_bytes_per_launch"] / rec["duration"] / 1e9
        out.append(rec)
    text = json.dumps(out, indent=1)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
