"""Extract the headline metrics of an `ncu --set full` report into one JSON line (profiles/*.json)."""
import csv, io, json, subprocess, sys
WANT = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct", "launch__registers_per_thread": "regs",
        "launch__grid_size": "grid", "launch__block_size": "block", "sm__cycles_elapsed.avg.per_second": "sm_ghz",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_lsu_pct"}
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {"report": path, "kernel": vals[hdr.index("Kernel Name")][:90]}
    for i, h in enumerate(hdr):
        if h in WANT:
            v = vals[i].replace(",", "")
            try: v = float(v)
            except ValueError: pass
            u = units[i]
            if WANT[h].startswith("dram_") and isinstance(v, float) and WANT[h] != "dram_pct":
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            if WANT[h] == "time_us" and isinstance(v, float):
                v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
            d[WANT[h]] = v
    if "dram_read" in d and "dram_write" in d: d["dram_traffic_bytes"] = d["dram_read"] + d["dram_write"]
    print(json.dumps(d))
if __name__ == "__main__":
    for p in sys.argv[1:]: main(p)
