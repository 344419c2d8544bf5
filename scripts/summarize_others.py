"""profiles/<tag>_other_kernels.md from gpurun_out/prof_others_<tag>.ncu-rep (ncu --set full of the non-MLP kernels)."""
import csv, collections, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_others_{tag}.ncu-rep", "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, rows = rr[0], rr[1], rr[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return 0.0
def to_bytes(r, k):
    u = units[ix[k]]
    return f(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def to_us(r):
    u = units[ix["gpu__time_duration.sum"]]
    return f(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6}.get(u, 1)
agg = collections.OrderedDict()
for r in rows:
    name = r[ix["Kernel Name"]].split("(")[0].replace("mp::", "")
    a = agg.setdefault(name, dict(n=0, us=0.0, dram=0.0, inst=0.0, lanes=0.0, issue=0.0, l1=0.0, l2=0.0, smt=0.0, big_us=0, big=None))
    us = to_us(r)
    a["n"] += 1
    a["us"] += us
    a["dram"] += to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    if us > a["big_us"]:
        a["big_us"], a["big"] = us, r
with open(f"profiles/{tag}_other_kernels.md", "w") as out:
    out.write(f"# ncu --set full of the non-MLP kernels of one warm step ({tag})\n\n"
              "Command: `MP_RENDER_STREAMS=0 ncu --set full --clock-control none -k regex:'deform|sampler|composite|...' -s 45 -c 45 "
              "python bench.py --steps 1 --warmup 1 --no-cpu-baseline` (configs[1]).  Times under ncu are serialised and cold-cache.  "
              "Per kernel: launches in the step, summed time, DRAM bytes and the resulting GB/s over the kernel's own time; the "
              "remaining columns are from its LONGEST launch (the data-carrying one; the sampler's late trips exit at once).\n\n"
              "| kernel | launches | us | DRAM MB | DRAM GB/s | longest us | DRAM % of peak | issue active % | lanes / instr | L1 hit % | L2 hit % | regs | achieved occupancy % |\n"
              "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        b = a["big"]
        out.write("| `%s` | %d | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.1f |\n" % (
            name, a["n"], a["us"], a["dram"] / 1e6, a["dram"] / 1e9 / (a["us"] * 1e-6) if a["us"] else 0, a["big_us"],
            f(b, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            f(b, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            f(b, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            f(b, "l1tex__t_sector_hit_rate.pct"), f(b, "lts__t_sector_hit_rate.pct"),
            int(f(b, "launch__registers_per_thread")), f(b, "sm__warps_active.avg.pct_of_peak_sustained_active")))
    out.write("\nReading: none of these kernels is near the HBM roof (the per-sample records they stream are a few tens of MB per "
              "step); they are latency / instruction bound integer-and-gather work, which is why DESIGN.md bounds the step by the "
              "tensor roofline of the MLP kernel and treats these as the overhead to hide or shrink.\n")
print(open(f"profiles/{tag}_other_kernels.md").read())
