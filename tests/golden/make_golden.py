"""Regenerate the golden fixtures from the UNMODIFIED reference (oracle/_ref/libbtbb_ref.so).

    make -C oracle ref && python tests/golden/make_golden.py

Inputs are deterministic (numpy PCG64 with fixed seeds, or this repo's own synthetic
capture generator); every OUTPUT in the fixtures comes from the reference's code.  The
reference builds its syndrome map once per process (bluetooth_packet.c:288), so each
btbb_init(k) variant is produced in its own subprocess.
"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import util  # noqa: E402
from util import B  # noqa: E402


def find_ac_stream(seed, n):
    """noise + planted sync words with 0..4 errors, some for LAP 0x9e8b33"""
    rng = np.random.default_rng(seed)
    s = rng.integers(0, 2, n + 64, dtype=np.uint8)
    util.plant_syncwords(s, rng, 150, 4)
    util.plant_syncwords(s, rng, 60, 6, laps=[0x9E8B33])
    return s


def synth_stream(ber):
    cfg = B.synth_cfg(600_000, stride=4096, ber=ber, seed=0xB200B7BB + int(ber * 1e4),
                      mix=("ID", "DM1", "DH1", "DM3", "FHS", "HV1", "DM5", "DH3"))
    return cfg, B.synth_host(cfg)


def gen_find(k_init):
    R = util.ref()
    assert R.btbb_init(k_init) == 0
    out = {"k_init": k_init, "cases": []}
    N = 1 << 20
    s = find_ac_stream(1234, N)
    for lap, ks in ((B.LAP_ANY, range(0, 6)), (0x9E8B33, (0, 1, 2, 5, 9))):
        for k in ks:
            h = util.find_all(R, "ref", s, N, lap, k)
            out["cases"].append({"stream": "rand1234", "n": N, "lap": lap, "k": k, "count": len(h),
                                 "sha256": util.digest(h), "head": h[:8].tobytes().hex()})
    cfg, s2 = synth_stream(0.005)
    n2 = len(s2) - 63
    for k in range(0, 5):
        h = util.find_all(R, "ref", s2, n2, B.LAP_ANY, k)
        out["cases"].append({"stream": "synth0.005", "n": n2, "lap": B.LAP_ANY, "k": k, "count": len(h),
                             "sha256": util.digest(h), "head": h[:8].tobytes().hex()})
    return out


def gen_primitives():
    R = util.ref()
    rng = np.random.default_rng(99)
    g = {}
    cws = [int(x) for x in rng.integers(0, 1 << 63, 256, dtype=np.uint64)] + [0xcc7b7268ff614e1b, 0xcc7d7268ff614e1b]
    g["syndrome"] = [[hex(c), hex(R.ref_gen_syndrome(c))] for c in cws]
    laps = [0, 0xffffff, 0x9e8b33] + [int(x) for x in rng.integers(0, 1 << 24, 200)]
    g["syncword"] = [[l, hex(R.btbb_gen_syncword(l))] for l in laps]
    g["barker_distance"] = [int(R.ref_barker_distance(b)) for b in range(128)]
    g["barker_correct"] = [hex(R.ref_barker_correct(b)) for b in range(128)]
    g["whitening"] = []
    for clk in range(64):
        bits = [int(R.ref_whitening_bit(int(R.ref_whitening_index(clk)) + i)) for i in range(127)]
        g["whitening"].append("".join(map(str, bits)))
    g["fec23"] = [int(R.ref_fec23(d)) for d in range(1024)]
    g["uap_from_hec"] = [[d, h, int(R.ref_uap_from_hec(d, h))] for d, h in
                         zip(rng.integers(0, 1024, 300).tolist(), rng.integers(0, 256, 300).tolist())]
    crc = []
    for _ in range(100):
        n = int(rng.integers(0, 300))
        bits = rng.integers(0, 2, max(n, 1), dtype=np.uint8)
        uap = int(rng.integers(0, 256))
        crc.append(["".join(map(str, bits[:n])), uap, int(R.ref_crcgen(bits.ctypes.data_as(C.c_char_p), n, uap))])
    g["crc"] = crc
    f13 = []
    for _ in range(100):
        L = int(rng.choice([18, 80, 7]))
        bits = np.repeat(rng.integers(0, 2, L, dtype=np.uint8), 3)
        flips = rng.random(3 * L) < rng.choice([0.0, 0.05, 0.12])
        bits ^= flips.astype(np.uint8)
        out = np.zeros(L, dtype=np.uint8)
        ok = R.ref_unfec13(bits.ctypes.data_as(C.c_char_p), out.ctypes.data_as(C.c_char_p), L)
        f13.append(["".join(map(str, bits)), L, int(ok), "".join(map(str, out))])
    g["unfec13"] = f13
    f23 = []
    for _ in range(300):
        L = int(rng.choice([1, 8, 10, 16, 25, 160]))
        nb = (L + 9) // 10
        sym = []
        for _b in range(nb):
            cw = int(R.ref_fec23(int(rng.integers(0, 1024))))
            for e in rng.choice(15, int(rng.choice([0, 0, 1, 1, 2])), replace=False):
                cw ^= 1 << int(e)
            sym += [(cw >> i) & 1 for i in range(15)]
        a = np.array(sym, dtype=np.uint8)
        out = np.zeros(nb * 10, dtype=np.uint8)
        ok = R.ref_unfec23(a.ctypes.data_as(C.c_char_p), L, out.ctypes.data_as(C.c_char_p))
        f23.append(["".join(map(str, sym)), L, int(ok), "".join(map(str, out)) if ok else ""])
    g["unfec23"] = f23
    g["sizeof_packet"] = int(R.ref_sizeof_packet())
    return g


def gen_decode():
    R = util.ref()
    out = {"streams": []}
    for ber in (0.0, 0.004, 0.02):
        cfg, s = synth_stream(ber)
        recs, tc, hp = [], [], []
        pl = util.planted_list(cfg)
        for p in pl:
            L = min(3125, len(s) - p.offset)
            recs.append(util.decode_one(R, "ref", s, p.offset, L, p.clk6, p.uap))
            hp.append(int(R.ref_header_present(s[p.offset:].ctypes.data, L)))
        for p in pl[:40]:
            L = min(3125, len(s) - p.offset)
            for c in range(64):
                tc.append(util.try_clock_one(R, "ref", s, p.offset, L, c))
        # truncated packets and a wrong clock / wrong UAP
        odd = []
        for p in pl[:60]:
            for L in (100, 121, 122, 130, 137, 200, 361, 362, 500):
                odd.append(util.decode_one(R, "ref", s, p.offset, min(L, len(s) - p.offset), p.clk6, p.uap))
            odd.append(util.decode_one(R, "ref", s, p.offset, 3125, (p.clk6 + 1) & 63, p.uap))
            odd.append(util.decode_one(R, "ref", s, p.offset, 3125, p.clk6, (p.uap + 1) & 255))
            odd.append(util.decode_one(R, "ref", s, p.offset, 3125, p.clk6, p.uap, 0))
        recs, tc, odd = np.array(recs), np.array(tc), np.array(odd)
        out["streams"].append({
            "ber": ber, "n_packets": len(pl), "decode_sha256": util.digest(recs),
            "rv_hist": {str(k): int(v) for k, v in zip(*np.unique(recs["rv"], return_counts=True))},
            "try_clock_sha256": util.digest(tc), "odd_sha256": util.digest(odd),
            "header_present": "".join(map(str, hp)),
            "decode_head": recs[:4].tobytes().hex()})
    return out


def gen_noise_types():
    """Every packet type, including EV3/EV4/EV5/HV2/DV/AUX1 which the synthetic transmitter
    does not build: try_clock + crc_check on random symbols lands on all 16 types."""
    R = util.ref()
    rng = np.random.default_rng(4242)
    recs = []
    for i in range(300):
        sym = rng.integers(0, 2, 3125, dtype=np.uint8)
        # make the header triplets agree so unfec13 succeeds
        hdr = np.repeat(rng.integers(0, 2, 18, dtype=np.uint8), 3)
        sym[68:122] = hdr
        L = int(rng.choice([3125, 1500, 700, 400, 250, 140]))
        for c in range(0, 64, 7):
            recs.append(util.try_clock_one(R, "ref", sym, 0, L, c))
        recs.append(util.decode_one(R, "ref", sym, 0, L, int(rng.integers(0, 64)), 0))
    recs = np.array(recs)
    return {"seed": 4242, "count": len(recs), "sha256": util.digest(recs),
            "type_hist": {str(k): int(v) for k, v in zip(*np.unique(recs["type"], return_counts=True))},
            "rv_hist": {str(k): int(v) for k, v in zip(*np.unique(recs["rv"], return_counts=True))}}


def gen_sieve():
    """Digests of the reference's btbb_uap_from_header results (oracle/_ref) on the sieve cases of
    tests/test_sieve.py -> tests/golden/sieve.json."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    import test_sieve
    R = util.ref()
    out = {}
    for name, kw in sorted(test_sieve.CASES.items()):
        stream, pkts, gs, laps, truth = util.sieve_case(**kw)
        st, rv = util.sieve_run(R, "ref", stream, pkts, gs)
        out[name] = {"shape": [len(pkts), len(gs) - 1], "sha256": [util.digest(st), util.digest(rv)],
                     "piconets_resolved": int(((st["flags"] >> 2) & 1).sum()),
                     "calls": int((rv >= 0).sum()), "returned_1": int((rv == 1).sum())}
    json.dump(out, open(os.path.join(HERE, "sieve.json"), "w"), indent=1)
    return out


def gen_pcap():
    """Digests of the files the reference's btbb_pcap_create_file / btbb_pcap_append_packet write for
    the cases of tests/test_pcap.py -> tests/golden/pcap.json."""
    import hashlib
    sys.path.insert(0, os.path.join(HERE, ".."))
    import test_pcap
    R = util.ref()
    R.ref_pcap_bredr.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_uint32, C.c_uint8, C.c_void_p]
    out = {}
    for name, kw in sorted(test_pcap.CASES.items()):
        stream, hits, pkts, meta = test_pcap.build_case(**kw)
        dec = test_pcap.oracle_records(stream, pkts)
        keep = test_pcap.well_defined(dec)
        for reflap, refuap in ((B.LAP_ANY, 0xFF), (0x9E8B33, 0x42)):
            h, p, m = hits[keep].copy(), pkts[keep].copy(), meta[keep].copy()
            rv = np.zeros(len(h), dtype=np.int32)
            path = os.path.join("/tmp", f"golden_{os.getpid()}.pcap").encode()
            assert R.ref_pcap_bredr(path, stream.ctypes.data, len(stream), h.ctypes.data, p.ctypes.data, m.ctypes.data,
                                    len(h), reflap, refuap, rv.ctypes.data) == 0
            data = open(path.decode(), "rb").read()
            os.remove(path.decode())
            out[f"{name}/{reflap:x}/{refuap:x}"] = {"shape": [int(keep.sum()), len(data)],
                                                   "sha256": hashlib.sha256(data).hexdigest(), "kept_of": len(keep)}
    json.dump(out, open(os.path.join(HERE, "pcap.json"), "w"), indent=1)
    return out


def gen_chain79():
    """The whole chain of BASELINE configs[2] through the reference (oracle/_ref): access codes of a
    79-channel capture, per hit btbb_decode with the true clock / UAP, the 64-clock try_clock + crc_check
    sweep, and btbb_uap_from_header per piconet -> tests/golden/chain79.json."""
    R = util.ref()
    assert R.btbb_init(2) == 0
    cfg, s, n = util.chain79_case()
    hits = util.find_all(R, "ref", s, n, B.LAP_ANY, 2)
    dec, sv, gs, laps = util.chain79_packets(cfg, hits)
    recs = np.array([util.decode_one(R, "ref", s, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"])) for p in dec])
    raw = np.array([util.decode_one_raw(R, "ref", s, int(p["offset"]), int(p["length"]), int(p["clkn"]), int(p["uap"])) for p in dec])
    tc = np.array([util.try_clock_one(R, "ref", s, int(p["offset"]), int(p["length"]), c) for p in dec for c in range(64)])
    st, rv = util.sieve_run(R, "ref", s, sv, gs)
    # the capture files the reference writes for these packets (pcap whole; pcapng: its enhanced packet blocks, pad bytes zeroed)
    import hashlib
    sys.path.insert(0, os.path.join(HERE, ".."))
    import test_pcap
    meta = util.chain79_meta(dec)
    R.ref_pcap_bredr.argtypes = [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_uint32, C.c_uint8, C.c_void_p]
    R.ref_pcapng_bredr.argtypes = R.ref_pcap_bredr.argtypes
    files = {}
    for kind, fn in (("pcap", R.ref_pcap_bredr), ("pcapng", R.ref_pcapng_bredr)):
        path = os.path.join("/tmp", f"golden_chain_{os.getpid()}.{kind}").encode()
        rvv = np.zeros(len(hits), dtype=np.int32)
        assert fn(path, s.ctypes.data, len(s), hits.ctypes.data, dec.ctypes.data, meta.ctypes.data, len(hits), B.LAP_ANY, 0xFF, rvv.ctypes.data) == 0
        files[kind] = open(path.decode(), "rb").read()
        os.remove(path.decode())
    png = test_pcap._epbs(files["pcapng"])
    out = {"pcap": [len(files["pcap"]), hashlib.sha256(files["pcap"]).hexdigest()],
           "pcapng_blocks": [len(png), hashlib.sha256(png).hexdigest()],
           "blocks": 6, "symbols": n, "hits": len(hits), "hits_sha256": util.digest(hits), "decode_sha256": util.digest(recs),
           "decode_raw_sha256": util.digest(raw), "try_clocks_sha256": util.digest(tc),
           "rv_hist": {str(k): int(v) for k, v in zip(*np.unique(recs["rv"], return_counts=True))},
           "piconets": len(gs) - 1, "piconets_resolved": int(((st["flags"] >> 2) & 1).sum()),
           "sieve_sha256": [util.digest(st), util.digest(rv)]}
    json.dump(out, open(os.path.join(HERE, "chain79.json"), "w"), indent=1)
    return out


def gen_hops():
    """Digests of windows of the reference's 2^27-entry hop table (gen_hop_pattern) for the addresses
    of tests/test_hops.py -> tests/golden/hops.json."""
    import hashlib
    sys.path.insert(0, os.path.join(HERE, ".."))
    import test_hops
    R = util.ref()
    out = {}
    for name, (addr, afh) in sorted(test_hops.CASES.items()):
        full = test_hops._seq(R, "ref", addr, afh, 0, 1 << 27)
        d = {"full": hashlib.sha256(full.tobytes()).hexdigest()}
        for first, n in test_hops.WINDOWS:
            d[f"{first}+{n}"] = hashlib.sha256(full[first:first + n].tobytes()).hexdigest()
        out[name] = d
    json.dump(out, open(os.path.join(HERE, "hops.json"), "w"), indent=1)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "hops":
        print(json.dumps(gen_hops()))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "chain79":
        print(json.dumps(gen_chain79()))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pcap":
        print(json.dumps(gen_pcap()))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sieve":
        print(json.dumps(gen_sieve()))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "find":
        print(json.dumps(gen_find(int(sys.argv[2]))))
        sys.exit(0)
    finds = []
    for k in (0, 1, 2, 3, 4):
        r = subprocess.run([sys.executable, __file__, "find", str(k)], capture_output=True, text=True, check=True)
        finds.append(json.loads(r.stdout))
    json.dump(finds, open(os.path.join(HERE, "find_ac.json"), "w"), indent=0)
    json.dump(gen_primitives(), open(os.path.join(HERE, "primitives.json"), "w"), indent=0)
    json.dump(gen_decode(), open(os.path.join(HERE, "decode.json"), "w"), indent=0)
    json.dump(gen_noise_types(), open(os.path.join(HERE, "noise_types.json"), "w"), indent=0)
    gen_sieve()
    subprocess.run([sys.executable, __file__, "chain79"], check=True, capture_output=True)
    gen_pcap()
    gen_hops()
    print("golden fixtures written to", HERE)
