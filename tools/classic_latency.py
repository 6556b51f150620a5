"""Per-call latency of the classic libbtbb calls: the product's default route (host small-call path
for single packets and short searches) beside the unmodified reference, through the same C harness
(tools/classic_latency.c, which dlopen()s the library it is given).

    python tools/classic_latency.py [--reps 20] [--find-ac] > profiles/r02_classic_latency.json

Without a GPU the product's btbb_init fails (by design), so --find-ac is for the GPU box; the
single-packet calls need no context and are timed anywhere.  The reference library travels to the GPU
box prebuilt (oracle/_ref); where it is absent only the product is timed."""
import argparse, ctypes as C, json, os, struct, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--packets", type=int, default=400)
    ap.add_argument("--find-ac", action="store_true")
    a = ap.parse_args()
    import util
    from util import B
    stride, lead = 6000, 500
    cfg = B.synth_cfg(stride * (a.packets + 1), stride=stride, ber=0.002, seed=777, mix=tuple(B.KIND))
    s = B.synth_host(cfg)
    rows = []
    kinds = {}
    for p in util.planted_list(cfg):
        if p.offset < lead or p.offset - lead + 4096 > len(s):
            continue
        win = s[p.offset - lead:p.offset - lead + 4096]
        rows.append(struct.pack("<4i", min(3125, 4096 - lead), p.clk6, p.uap, lead) + win.tobytes())
        kinds[p.kind] = kinds.get(p.kind, 0) + 1
    names = {v: k for k, v in B.KIND.items()}
    tmp = tempfile.mkdtemp()
    pkf = os.path.join(tmp, "packets.bin")
    with open(pkf, "wb") as f:
        f.write(struct.pack("<i", len(rows)))
        for r in rows:
            f.write(r)
    exe = os.path.join(tmp, "classic_latency")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "classic_latency.c"), "-ldl"], check=True)
    libs = [("product", B.LIB_PATH)]
    if util.have_ref():
        libs.append(("reference", util.REF_SO))
    out = {"packets": len(rows), "mix": {names.get(k, str(k)): v for k, v in sorted(kinds.items())},
           "reps": a.reps, "unit": "microseconds per call, mean", "host": {"cores": os.cpu_count()}}
    for name, path in libs:
        cmd = [exe, path, pkf, str(a.reps)] + (["find_ac"] if a.find_ac else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            out[name] = {"error": r.stderr.strip()[-300:]}
            continue
        out[name] = json.loads(r.stdout.strip().splitlines()[-1])
    if "reference" in out and "checksums" in out.get("product", {}) and "checksums" in out["reference"]:
        p, q = out["product"], out["reference"]
        out["same_results"] = p["checksums"][:2] == q["checksums"][:2] and (not a.find_ac or p["checksums"][2] == q["checksums"][2])
        out["ratio_reference_over_product"] = {k: (q[k] / p[k] if p[k] > 0 and q[k] > 0 else None)
                                                for k in ("decode_us", "uap_sweep_64_clocks_us", "find_ac_4096_us")}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
