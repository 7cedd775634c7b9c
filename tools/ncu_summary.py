"""Summarise an .ncu-rep (`ncu --set full`) into a small JSON for profiles/: per launch the metrics DESIGN.md quotes,
the stall-reason totals of the sampled warps and the total warp instructions.   python tools/ncu_summary.py x.ncu-rep out.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
ix = {h: i for i, h in enumerate(hdr)}
launches = []
for r in rows[2:]:
    d = {"kernel": r[ix["Kernel Name"]][:90]}
    for w in WANT:
        if w in ix:
            d[w] = r[ix[w]] + " " + units[ix[w]]
    launches.append(d)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
        continue
    if cur is not None:
        cur.append(r)
# the source page repeats every launch twice (two views); keep every other block
for li, b in enumerate(blocks[::max(1, len(blocks) // max(1, len(launches)))][:len(launches)]):
    h = {n: i for i, n in enumerate(b[0])}
    st = [n for n in b[0] if n.startswith("stall_") and "Not Issued" not in n]
    tot = {n[6:]: sum(int(r[h[n]]) for r in b[1:]) for n in st}
    launches[li]["warp_stall_samples"] = {k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0}
    launches[li]["sass_instructions"] = len(b) - 1
json.dump({"report": rep.split("/")[-1], "launches": launches}, open(out, "w"), indent=1)
print(json.dumps(launches, indent=1)[:3000])
