"""Summarise an .ncu-rep (one or more captured launches) into a small JSON for profiles/:
    python tools/ncu_summary.py report.ncu-rep out.json
Keys follow /opt/skills/guides/B200_PROFILING.md: duration, DRAM bytes, L2 / L1 bytes, pipe utilisation, stall mix."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {
    "gpu__time_duration.sum": "duration",
    "launch__grid_size": "grid", "launch__block_size": "block", "launch__cluster_size": "cluster_size",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
    "lts__t_bytes.sum": "l2_bytes", "l1tex__t_bytes.sum": "l1_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct_active",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "sm__cycles_active.avg": "sm_cycles_active_avg",
}
stall = "smsp__average_warps_issue_stalled_"
launches = []
for r in rows[2:]:
    d = {}
    name_i = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    if name_i is not None:
        d["kernel"] = r[name_i]
    stalls = {}
    for i, h in enumerate(hdr):
        if h in want:
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            d[want[h]] = v
            d[want[h] + "_unit"] = units[i]
        elif h.startswith(stall) and h.endswith("_per_issue_active.ratio"):
            try:
                stalls[h[len(stall):-len("_per_issue_active.ratio")]] = round(float(r[i]), 3)
            except ValueError:
                pass
    d["stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    launches.append(d)


def to_bytes(v, unit):
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


first = launches[0]
summary = {"report": rep.split("/")[-1], "launches": launches}
if "dram_bytes_read" in first:
    summary["dram_bytes_per_launch"] = int(to_bytes(first["dram_bytes_read"], first["dram_bytes_read_unit"]) +
                                          to_bytes(first["dram_bytes_write"], first["dram_bytes_write_unit"]))
json.dump(summary, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in first.items() if not k.endswith("_unit")}, indent=1)[:1500])
