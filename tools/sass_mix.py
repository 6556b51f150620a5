"""Static evidence per kernel from the built objects (no GPU): resource usage (cuobjdump -res-usage)
and the SASS opcode mix (cuobjdump -sass), written as profiles/<tag>_sass_summary.txt.

    python tools/sass_mix.py [tag]
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "libbtbb_b200", "build")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
WANT = ["scan_promisc_v7ILi0ELi5ELi1ELi0ELb0", "scan_promisc_v7ILi0ELi5ELi1ELi0ELb1", "scan_promisc_v7ILi0ELi5ELi0ELi1ELb0",
        "scan_promisc_v7ILi0ELi5ELi0ELi2ELb0", "scan_known_v4ILb0ELb0ELb0ELb0", "slab_sort_kernel", "slab_scan_kernel", "scan_prep_kernel",
        "decode_kernelILi0ELi8ELi32", "decode_kernelILi1ELi24ELi16", "decode_kernelILi2ELi8ELi32", "sieve_kernel", "hop_sequence_kernel",
        "hop_winnow_kernel", "capture_write_kernel"]
PIPE = [("ALU", r"^(LOP3|SHF|IADD3|IADD|LEA|PRMT|ISETP|SEL|IMNMX|VIADD|VIMNMX|BMSK|SGXT|MOV|PLOP3|ICMP|ISCADD)"),
        ("FMA (IMAD etc.)", r"^(IMAD|FFMA|FMUL|FADD|IDP)"), ("XU", r"^(FLO|POPC|BREV|MUFU|I2F|F2I)"),
        ("LSU shared", r"^(LDS|STS|ATOMS|LDSM)"), ("LSU global", r"^(LDG|STG|ATOMG|RED|ATOM|LD\.|ST\.|LDL|STL|CCTL)"),
        ("bulk copy / TMA", r"^(UBLKCP|UTMA|SYNCS|UBLKPF)"), ("warp (SHFL/VOTE/REDUX/MATCH)", r"^(SHFL|VOTE|REDUX|MATCH|WARPSYNC|VOTEU)"),
        ("control", r"^(BRA|BSSY|BSYNC|EXIT|RET|CALL|BAR|NOP|BRX|JMP|YIELD|DEPBAR|ERRBAR|MEMBAR|FENCE)"),
        ("uniform", r"^(U[A-Z0-9]+|R2UR|S2UR|S2R|CS2R|LDC|ULDC|R2P|P2R)")]

out = []
for obj in sorted(os.listdir(OBJ)):
    if not obj.endswith(".cu.o"):
        continue
    path = os.path.join(OBJ, obj)
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = line.strip()
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, ops = None, collections.defaultdict(list)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            ops[cur].append(m.group(1))
    for fn, lst in ops.items():
        key = next((w for w in WANT if w in fn), None)
        if not key:
            continue
        demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
        demangled = re.sub(r"\(anonymous namespace\)::|_GLOBAL__N__\w+::", "", demangled).split("(")[0].replace("void ", "")
        hist = collections.Counter(o.split(".")[0] for o in lst)
        cls = collections.Counter()
        for o in lst:
            for name, rx in PIPE:
                if re.match(rx, o):
                    cls[name] += 1
                    break
            else:
                cls["other"] += 1
        special = sorted({o for o in lst if re.match(r"^(LDG\.E\.\S*(128|256)|LDG\.E\.\S*NA|UBLKCP|UTMA|REDUX|IDP\.4A|IDP|FLO|BMSK|STG\.E\.128|LDS\.128|STS\.128|MATCH)", o)})
        out.append((obj, demangled, usage.get(fn, ""), len(lst), cls, hist.most_common(12), special))

with open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt"), "w") as f:
    f.write("Static view of the shipped kernels (cuobjdump -res-usage / -sass of libbtbb_b200/build/*.cu.o, sm_100a):\n"
            "instruction counts are whole-kernel SASS lines, not executed counts -- they show which pipes a kernel's\n"
            "code is made of and that the wide loads / bulk copies / warp reductions named in DESIGN.md are really there.\n\n")
    for obj, name, use, n, cls, top, special in sorted(out, key=lambda x: (x[0], x[1])):
        f.write(f"{name}   [{obj}]\n  {use}\n  {n} SASS instructions: " + ", ".join(f"{k} {v}" for k, v in cls.most_common()) + "\n")
        f.write("  top opcodes: " + ", ".join(f"{k} {v}" for k, v in top) + "\n")
        f.write("  notable: " + (", ".join(special) if special else "-") + "\n\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")).read()[:6000])
