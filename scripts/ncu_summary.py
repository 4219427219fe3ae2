"""Summarise .ncu-rep captures into one JSON: python scripts/ncu_summary.py out.json rep1.ncu-rep [rep2 ...]
Per kernel launch: duration, DRAM bytes / throughput, L2 hit rate, tensor-pipe and issue utilisation, registers, occupancy."""
import csv, io, json, subprocess, sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_B",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
}
out = {}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rec = {"kernel": d.get("Kernel Name", "?")[:90]}
        for k, name in WANT.items():
            if k in d and d[k] not in ("", "n/a"):
                try:
                    v = float(d[k].replace(",", ""))
                except ValueError:
                    continue
                unit = u.get(k, "")
                if name == "duration_us":
                    v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
                if name.endswith("_MB"):
                    v = {"byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(unit, v)
                rec[name] = round(v, 3)
        if "dram_read_MB" in rec and "duration_us" in rec:
            rec["dram_GBs"] = round((rec.get("dram_read_MB", 0) + rec.get("dram_write_MB", 0)) / rec["duration_us"] * 1e3, 1)
        out.setdefault(rep.split("/")[-1].replace(".ncu-rep", ""), []).append(rec)
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1)[:6000])
