"""Turn the raw outputs of tools/gpu_artifacts.sh (gpurun_out/art/) into the tracked summaries
under profiles/ (round tag as argv[1], default r01)."""
import csv
import json
import os
import shutil
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ART = os.path.join(ROOT, "gpurun_out", "art")
PRO = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

for src, dst in (("bench_n1.json", f"{tag}_bench_n1.json"), ("bench_reference_arm.json", f"{tag}_bench_reference_arm.json"),
                 ("chain_config3.json", f"{tag}_chain_config3.json"), ("sweep_config5.json", f"{tag}_sweep_config5.json"),
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
        name = r[ik].split("(")[0].replace("(anonymous namespace)::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e6
    tot = sum(v[1] for k, v in agg.items() if not k.startswith("synth_"))
    with open(os.path.join(PRO, f"{tag}_launch_list_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu\n")
        f.write("(cold-cache, serialised launches: compare SHARES; 10^10 symbols, promiscuous k=2; 7 find_ac passes = 3 warm-up + 2 steps + 2 kernel-only)\n\n")
        f.write(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s}  share of non-synth time\n")
        for k, (n, ms) in agg.items():
            share = "" if k.startswith("synth_") else f"{100 * ms / tot:6.2f} %"
            f.write(f"{k[:60]:60s} {n:8d} {ms:10.3f}  {share:>22s}\n")

# ---- key metrics of the full capture ----
p = os.path.join(ART, "scan_v7_full_raw.csv")
if os.path.exists(p) and os.path.getsize(p) > 0:
    rows = list(csv.reader(open(p)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    get = lambda k: vals[hdr.index(k)]
    KEYS = ["dram__bytes_read.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
            "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    name = get("Kernel Name") if "Kernel Name" in hdr else "scan_promisc_v7"
    with open(os.path.join(PRO, f"{tag}_scan_v7_ncu_full_key_metrics.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on -k regex:scan_promisc_v7 -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu\n")
        f.write(f"kernel {name} (shipped), grid {get('Grid Size') if 'Grid Size' in hdr else '?'} x block {get('Block Size') if 'Block Size' in hdr else '?'}, 10^10 symbols (B200)\n\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k:90s} {units[i]:12s} {vals[i]}\n")
        f.write("\nwarp stall reasons (warps per issue-active cycle):\n")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h:
                f.write(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {float(vals[i]):.3f}\n")
    rd = float(get("dram__bytes_read.sum")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(get("dram__bytes_write.sum")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
    bench = json.load(open(os.path.join(PRO, f"{tag}_bench_n1.json")))
    alg = bench["roofline"]["algorithmic_bytes_per_launch"]
    json.dump({"kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "algorithmic_bytes_per_launch": alg,
               "note": f"ncu --set full, 10^10-symbol launch inside bench.py (profiles/{tag}_scan_v7_ncu_full_key_metrics.txt); "
                       f"traffic / algorithmic bytes = {(rd + wr) / alg:.4f}"},
              open(os.path.join(PRO, "traffic.json"), "w"), indent=1)
print("profiles updated")
