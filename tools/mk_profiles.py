"""Turn the raw outputs of tools/gpu_artifacts.sh (gpurun_out/art/) into the tracked summaries
under profiles/ (round tag as argv[1], default r02)."""
import csv
import json
import os
import shutil
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ART = os.path.join(ROOT, "gpurun_out", "art")
PRO = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"

for src, dst in (("bench_n1.json", f"{tag}_bench_n1.json"), ("bench_reference_arm.json", f"{tag}_bench_reference_arm.json"),
                 ("launches.csv", f"{tag}_launches.csv")):
    p = os.path.join(ART, src)
    if os.path.exists(p) and os.path.getsize(p) > 0:
        shutil.copy(p, os.path.join(PRO, dst))

# ---- launch list summary ----
p = os.path.join(ART, "launches.csv")
if os.path.exists(p):
    rows = [r for r in csv.reader(l for l in open(p) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e6
    with open(os.path.join(PRO, f"{tag}_launch_list_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 600 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu\n")
        f.write("(cold-cache, serialised launches: compare SHARES within a group, not absolutes.  The run is bench.py's whole N=1 line:\n"
                " 5 promiscuous scans over 10^10 symbols (3 warm-up + 2 timed steps), the known-LAP block, the config-3 chain block.)\n\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total ms':>10s}\n")
        for k, (n, ms) in agg.items():
            f.write(f"{k[:72]:72s} {n:8d} {ms:10.3f}\n")
        # the timed workload's steps: every 10^10-symbol launch of the bulk kernel plus the launches that follow it up to the next scan
        seq = [(r[ik].split("(")[0], float(r[iv].replace(",", "")) / 1e6) for r in rows[1:]]
        steps, cur = [], None
        for name, ms in seq:
            if "scan_promisc_v7" in name or "scan_known_v4" in name:
                if cur:
                    steps.append(cur)
                cur = [ms, 0.0] if ("scan_promisc_v7" in name and ms > 1.5) else None
            elif cur and any(t in name for t in ("scan_promisc_kernel", "slab_scan_kernel", "slab_sort_kernel")):
                cur[1] += ms
        if cur:
            steps.append(cur)
        if steps:
            k_ms = sum(x[0] for x in steps) / len(steps)
            o_ms = sum(x[1] for x in steps) / len(steps)
            f.write(f"\nper 10^10-symbol step ({len(steps)} of them): bulk scan kernel {k_ms:.3f} ms, tile kernel on the tail + slab_scan + slab_sort {o_ms:.3f} ms "
                    f"-> the bulk kernel's share of the step's kernel time is {100 * k_ms / (k_ms + o_ms):.1f} %\n"
                    f"(bench.py, CUDA events, same command without ncu: kernel_ms / ms_per_step in {tag}_bench_n1.json)\n")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
WORK = {"v7": "10^10 symbols, promiscuous, tables for 2 errors (BASELINE configs[1])", "known": "4 x 10^9 symbols, known LAP, k = 2",
        "k3": "4 x 10^9 symbols, promiscuous, tables for 3 errors", "k4": "4 x 10^9 symbols, tables for 4 errors", "k5": "2 x 10^9 symbols, tables for 5 errors",
        "decode1": "79 000 packets of the config-3 capture, 64-clock sweep, 64 full records per packet", "decode0": "79 000 packets, btbb_decode with the true clock / UAP",
        "tc16": "79 000 packets, 64-clock sweep, compact 16-bit results (the UAP sieve's input)", "sieve": "UAP sieve rounds of one call (79 000 packets, 75 piconets): the longest launch",
        "hops": "2^27-entry hop sequence of one address", "winnow": "hop reversal: 2^21 candidates x 12 observations", "slabsort": "ordering pass of a 10^10-symbol scan (10^6 hits)",
        "capture": "pcap records of 158 005 packets (payload 0..183 bytes) formatted on the device: the write kernel"}
summary = []
for name in ("v7", "known", "k3", "k4", "k5", "decode1", "decode0", "tc16", "sieve", "hops", "winnow", "slabsort", "capture"):
    p = os.path.join(ART, f"ncu_{name}_raw.csv")
    if not (os.path.exists(p) and os.path.getsize(p) > 0):
        continue
    rows = list(csv.reader(open(p)))
    hdr, units = rows[0], rows[1]
    # several launches captured (sieve rounds): keep the longest
    it = hdr.index("gpu__time_duration.sum")
    def dur(r):
        v = float(r[it].replace(",", ""))
        return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[it], 1.0)
    vals = max(rows[2:], key=dur)
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name
    with open(os.path.join(PRO, f"{tag}_ncu_{name}.txt"), "w") as f:
        f.write(f"ncu --set full --clock-control none (B200, sm_100a) -- {kname}\nworkload: {WORK[name]}\n"
                f"(a profiled launch: use the counters, not the duration, as evidence; timings are in {tag}_bench_n1.json)\n\n")
        for k in KEYS:
            if k in hdr:
                f.write(f"{k:95s} {units[hdr.index(k)]:14s} {vals[hdr.index(k)]}\n")
    get = lambda k: vals[hdr.index(k)] if k in hdr else ""
    summary.append((name, kname.split("(")[0][-60:], dur(vals), get("dram__bytes_read.sum") + " " + units[hdr.index("dram__bytes_read.sum")],
                    get("dram__bytes_write.sum") + " " + units[hdr.index("dram__bytes_write.sum")],
                    get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                    get("sm__warps_active.avg.pct_of_peak_sustained_active"), get("launch__registers_per_thread")))
    if name == "v7":
        rd = float(get("dram__bytes_read.sum").replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(get("dram__bytes_write.sum").replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
        json.dump({"dram_bytes_per_launch": rd + wr, "note": f"ncu --set full, 10^10-symbol launch of scan_promisc_v7<0,5,1> (profiles/{tag}_ncu_v7.txt); "
                   f"traffic / algorithmic bytes = {(rd + wr) / 10016002047:.4f}"}, open(os.path.join(PRO, "traffic.json"), "w"))
with open(os.path.join(PRO, f"{tag}_ncu_summary.txt"), "w") as f:
    f.write("one ncu --set full capture per kernel on the path (key counters; full key lists in the per-kernel files)\n\n")
    f.write(f"{'capture':9s} {'ms':>8s} {'dram read':>16s} {'dram write':>16s} {'dram%':>6s} {'issue%':>7s} {'alu%':>6s} {'smem%':>6s} {'warps%':>7s} {'regs':>5s}  kernel\n")
    for s in summary:
        f.write(f"{s[0]:9s} {s[2]:8.3f} {s[3]:>16s} {s[4]:>16s} {float(s[5] or 0):6.1f} {float(s[6] or 0):7.1f} {float(s[7] or 0):6.1f} {float(s[8] or 0):6.1f} {float(s[9] or 0):7.1f} {s[10]:>5s}  {s[1]}\n")
print(open(os.path.join(PRO, f"{tag}_ncu_summary.txt")).read())
print(open(os.path.join(PRO, f"{tag}_launch_list_summary.txt")).read())
