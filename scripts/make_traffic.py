"""profiles/traffic.json from an `ncu --set full` capture of one bench pass (scripts/gpu_final_profile.sh <tag> b):
   python scripts/make_traffic.py gpurun_out/prof_<tag>.ncu-rep <tag>
Per-launch DRAM bytes and warp-instruction counts of every hot kernel; bench.py reads them for `roofline.traffic`
and `issue_slot_frac` (counters cannot be read outside a profiler, so they come from the committed capture)."""
import csv, json, os, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
idx = {h: i for i, h in enumerate(rows[0])}
units = rows[1]
names = {"generate_kernel": "generate_maps", "gmm_sample_kernel": "gmm_sample", "dda_kernel": "dda_gridcheck",
         "verdict_kernel": "verdict_fused", "compact_bits": "compact_survivors"}


def val(r, k):
    v, u = float(r[idx[k]]), units[idx[k]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


out = {}
for r in rows[2:]:
    key = next((v for k, v in names.items() if k in r[idx["Kernel Name"]]), None)
    if key is None:
        continue
    d = out.setdefault(key, {"dram_bytes_per_launch": 0, "warp_instructions_per_launch": 0, "ncu_time_us": 0.0, "launches_summed": 0})
    if key != "compact_survivors" and d["launches_summed"]:
        continue                                            # first launch of each kernel; the compaction is two kernels
    d["dram_bytes_per_launch"] += int(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
    d["warp_instructions_per_launch"] += int(float(r[idx["smsp__inst_executed.sum"]]))
    d["ncu_time_us"] += float(r[idx["gpu__time_duration.sum"]])
    d["launches_summed"] += 1
    if key != "compact_survivors":
        d.update(issue_active_pct=float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                 lanes_per_instruction=float(r[idx["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                 smem_wavefronts=int(float(r[idx["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]])),
                 registers=int(float(r[idx["launch__registers_per_thread"]])),
                 warps_active_pct=float(r[idx["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
                 fp64_pipe_pct=float(r[idx["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]]))
out["_source"] = "ncu --set full --clock-control none, capture %s (profiles/%s_ncu_full_summary.csv)" % (tag, tag)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
