"""BASELINE configs[4]: max_ac_errors 0..4 x injected BER sweep -- Gbit/s and detection rate.

    python tools/sweep.py [--symbols N] [--check-symbols M] > profiles/r01_sweep.json

Detection rate = planted access codes reported at their exact offset with the right LAP /
planted.  For every (k, BER) cell the GPU hit list over the first M symbols is also compared
byte-for-byte with the unmodified reference (oracle/_ref, one subprocess per btbb_init(k))
or, where that was not built, with the oracle port."""
import argparse, ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

BERS = [0.0, 0.001, 0.005, 0.01, 0.02, 0.05]
MIX = ("ID", "DM1", "DM3", "DH1", "FHS")


def cpu_cell(k, ber, m):
    """hit-list digest of the CPU checker for one cell (run in a subprocess: table built once)."""
    import util
    from util import B
    cfg = B.synth_cfg(m + 63, stride=10000, ber=ber, mix=MIX)
    s = B.synth_host(cfg)
    if util.have_ref():
        L = util.ref(); assert L.btbb_init(k) == 0
        h = util.find_all(L, "ref", s, m, B.LAP_ANY, k)
        kind = "reference"
    else:
        L = util.oracle(); assert L.orc_init(k) == 0
        h = util.find_all(L, "orc", s, m, B.LAP_ANY, k)
        kind = "port"
    return {"kind": kind, "count": len(h), "sha": util.digest(h)}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cpu":
        k, m = int(sys.argv[2]), int(sys.argv[3])
        print(json.dumps([cpu_cell(k, b, m) for b in BERS]))
        sys.exit(0)
    ap = argparse.ArgumentParser()
    ap.add_argument("--symbols", type=int, default=2 * 10**9)
    ap.add_argument("--check-symbols", type=int, default=1 << 23)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    import torch, util
    from util import B
    lib = B.lib()
    n, m = a.symbols, a.check_symbols
    d = torch.empty(n + 63, dtype=torch.uint8, device="cuda")
    cap = n // 10000 * 2 + (1 << 20)
    d_hits = torch.zeros((cap, 16), dtype=torch.uint8, device="cuda")
    out = []
    for k in range(5):
        cpu = json.loads(subprocess.run([sys.executable, __file__, "cpu", str(k), str(m)], capture_output=True,
                                        text=True, check=True).stdout.strip().splitlines()[-1])
        ctx = B.Context(0, k)
        for bi, ber in enumerate(BERS):
            cfg = B.synth_cfg(n + 63, stride=10000, ber=ber, mix=MIX)
            B.check(lib.btbb_b200_synth_dev(C.byref(cfg), d.data_ptr(), 0)); torch.cuda.synchronize()
            cnt, rc = ctx.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, k=k)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.iters)]
            for e0, e1 in ev:
                e0.record(); ctx.find_ac_dev(d.data_ptr(), n, d_hits.data_ptr(), cap, k=k); e1.record()
            torch.cuda.synchronize()
            ms = sorted(x.elapsed_time(y) for x, y in ev)[len(ev) // 2]
            hits = d_hits[:cnt].cpu().numpy().reshape(-1).view(B.HIT_DTYPE)
            # detection rate against the planted ground truth (sampled slots)
            offs = hits["offset"]
            slots = np.linspace(0, n // 10000 - 2, 4000).astype(np.int64)
            found = 0
            for sl in slots:
                p = B.planted(cfg, int(sl))
                i = np.searchsorted(offs, p.offset)
                found += bool(i < cnt and offs[i] == p.offset and hits["lap"][i] == p.lap)
            # exactness on the prefix
            c2, _ = ctx.find_ac_dev(d.data_ptr(), m, d_hits.data_ptr(), cap, k=k)
            pre = d_hits[:c2].cpu().numpy().reshape(-1).view(B.HIT_DTYPE)
            same = (len(pre) == cpu[bi]["count"]) and (util.digest(pre) == cpu[bi]["sha"])
            out.append({"k": k, "ber": ber, "gbit_s": n / (ms / 1e3) / 1e9, "ms": ms, "hits": int(cnt),
                        "detection_rate": found / len(slots), "matches_cpu": bool(same), "cpu_kind": cpu[bi]["kind"],
                        "prefix_hits": int(c2)})
            print(json.dumps(out[-1]), flush=True)
        ctx.close()
    print(json.dumps({"summary": "all_match", "value": all(o["matches_cpu"] for o in out)}))
