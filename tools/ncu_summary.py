"""ncu report (.ncu-rep) -> compact JSON summary for profiles/: python tools/ncu_summary.py in.ncu-rep out.json"""
import csv, io, json, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__cluster_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    rec = {"ID": r[idx["ID"]], "Kernel Name": r[idx["Kernel Name"]].replace("mra::<unnamed>::", "")[:110],
           "Grid Size": r[idx["Grid Size"]], "Block Size": r[idx["Block Size"]], "units": {}}
    for w in WANT:
        if w in idx:
            rec[w] = r[idx[w]]
            rec["units"][w] = units[idx[w]]
    def to_bytes(name):
        v, u = float(r[idx[name]]), units[idx[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    rec["dram_bytes"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    out.append(rec)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(len(out), "kernels ->", sys.argv[2])
