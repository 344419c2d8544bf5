"""Turn gpurun_out ncu outputs into committed summaries under profiles/.
    python scripts/summarize_profile.py r1
"""
import csv, collections, re, subprocess, sys, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = "profiles"
os.makedirs(out_dir, exist_ok=True)

# ---- launch list --------------------------------------------------------------------------
rows = []
with open(f"gpurun_out/launches_{tag}.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((int(r["ID"]), r["Kernel Name"].split("(")[0], float(r["Metric Value"].replace(",", ""))))
idx = [i for i, r in enumerate(rows) if "camera_rays" in r[1]]
s = idx[1]                       # second resident-input step (warm)
e = next(i for i in range(s, len(rows)) if "final_compose" in rows[i][1]) + 1
agg = collections.OrderedDict()
tot = 0.0
for r in rows[s:e]:
    agg.setdefault(r[1], [0, 0.0])
    agg[r[1]][0] += 1
    agg[r[1]][1] += r[2]
    tot += r[2]
with open(f"{out_dir}/{tag}_launches.md", "w") as f:
    f.write(f"# Launch list of one warm step ({tag})\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 "
            "--warmup 1 --no-cpu-baseline` (configs[1]: 4096 rays, 2 persons, S=128).  Per-launch times under ncu are "
            "serialised and cold-cache: compare shares, not absolutes.\n\n")
    f.write(f"{e - s} launches, {tot / 1e6:.3f} ms of device time\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f} % |\n")
    f.write("\nPer-launch durations of `tc_chain_kernel` in that step (us): " +
            ", ".join(f"{r[2] / 1e3:.1f}" for r in rows[s:e] if "tc_chain" in r[1]) + "\n")
print(open(f"{out_dir}/{tag}_launches.md").read())

# ---- full capture of the dominant kernel ---------------------------------------------------
rep = f"gpurun_out/prof_tc_{tag}.ncu-rep"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
with open(f"{out_dir}/{tag}_tc_kernel.md", "w") as f:
    f.write(f"# ncu --set full capture of `tc_chain_kernel` ({tag})\n\n")
    f.write("Command: `ncu --set full --clock-control none --import-source on -k regex:tc_chain -s 13 -c 13 python bench.py "
            "--steps 1 --warmup 1 --no-cpu-baseline --no-extras`.  Launch order per step: person 0: sdf-only x5 "
            "(sampler trips; inactive ones exit), shade; person 1: same; the background chain last (ids 0-5 person 0, 6-11 person 1, 12 background).\n\n")
    for r in rr[2:]:
        t = r[ix["gpu__time_duration.sum"]]
        f.write(f"## launch id {r[ix['ID']]}  ({t} {units[ix['gpu__time_duration.sum']]})\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in want:
            if w in ix:
                f.write(f"| {w} | {r[ix[w]]} | {units[ix[w]]} |\n")
        f.write("\n")
    # DRAM traffic of the kernel over one step (the 8 captured launches = every non-trivial launch of a step;
    # the 5 skipped leading launches are person 0's sampler trips: one 0.1 ms launch + 4 empty ones)
    import json
    tr = 0.0
    for r in rr[2:]:
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tr += float(r[ix["dram__bytes_read.sum"]]) * mult[units[ix["dram__bytes_read.sum"]]]
        tr += float(r[ix["dram__bytes_write.sum"]]) * mult[units[ix["dram__bytes_write.sum"]]]
    json.dump({"tc_chain_kernel_dram_bytes_per_step": tr, "launches_captured": len(rr) - 2,
               "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, profiles/%s_tc_kernel.md" % tag},
              open(f"{out_dir}/{tag}_traffic.json", "w"), indent=1)
    f.write(f"DRAM traffic summed over these launches (one step): {tr / 1e9:.2f} GB\n\n")
    # stall reasons of the first big launch (source page)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:tc_chain:1"],
                         capture_output=True, text=True).stdout
    sr = list(csv.reader(src.splitlines()))
    h2 = sr[1]
    i2 = {h: i for i, h in enumerate(h2)}
    sc = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
    totc = collections.Counter()
    ops = collections.Counter()
    for r in sr[2:]:
        if len(r) < len(h2):
            continue
        try:
            ie = int(r[i2["Instructions Executed"]])
        except Exception:
            continue
        for c in sc:
            totc[c] += int(r[i2[c]] or 0)
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[i2["Source"]].strip())
        ops[(m.group(2).split(".")[0] if m else "?")] += ie
    S = sum(totc.values()) or 1
    f.write("## warp stall samples, first shade launch (all warps incl. the spinning loader / MMA warps)\n\n| reason | share |\n|---|---:|\n")
    for c, v in totc.most_common(8):
        f.write(f"| {c} | {100 * v / S:.1f} % |\n")
    T = sum(ops.values()) or 1
    f.write("\n## instruction mix (warp instructions executed)\n\n| opcode | share |\n|---|---:|\n")
    for k, v in ops.most_common(14):
        f.write(f"| {k} | {100 * v / T:.1f} % |\n")
print(open(f"{out_dir}/{tag}_tc_kernel.md").read()[:6000])
