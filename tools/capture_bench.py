"""Capture-file bytes for a batch of decoded packets: records copied to the host and formatted there
(btbb_b200_pcap_bredr_records) against formatted on the device and copied out as file bytes
(btbb_b200_capture_records_dev).  Prints one JSON object.

    python tools/capture_bench.py [--packets N]
"""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--packets", type=int, default=158005)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    import torch
    from libbtbb_b200 import binding as B
    L = B.lib()
    n = a.packets
    rng = np.random.default_rng(5)
    hits = np.zeros(n, dtype=B.HIT_DTYPE)
    hits["offset"] = np.arange(n) * 4096
    hits["lap"] = rng.integers(0, 1 << 24, n)
    rec = np.zeros(n, dtype=B.DECODED_DTYPE)
    rec["payload_length"] = rng.choice([0, 17, 20, 27, 121, 183], n)      # ID / DM1 / FHS / DH1 / DM3 / DH3 bodies
    rec["payload"] = rng.integers(0, 256, (n, 344), dtype=np.uint8)
    rec["header_packed"] = rng.integers(0, 1 << 18, n)
    meta = np.zeros(n, dtype=B.PCAP_META_DTYPE)
    meta["ns"] = np.arange(n, dtype=np.uint64) * 625000
    meta["sigdbm"], meta["noisedbm"] = -40, -90
    ctx = B.Context(0, 2)
    dh = torch.from_numpy(hits.view(np.uint8).copy()).cuda()
    dr = torch.from_numpy(rec.view(np.uint8).copy()).cuda()
    dm = torch.from_numpy(meta.view(np.uint8).copy()).cuda()
    h_hits = torch.empty(dh.shape, dtype=torch.uint8, pin_memory=True)
    h_rec = torch.empty(dr.shape, dtype=torch.uint8, pin_memory=True)
    out = {"packets": n}
    for fmt, name in ((0, "pcap"), (1, "pcapng")):
        host_fn = L.btbb_b200_pcap_bredr_records if fmt == 0 else L.btbb_b200_pcapng_bredr_blocks
        need = host_fn(hits.ctypes.data, rec.ctypes.data, meta.ctypes.data, n, B.LAP_ANY, 0xFF, None, 0)
        buf = np.zeros(need, dtype=np.uint8)
        d_out = torch.empty(need, dtype=torch.uint8, device="cuda")
        h_out = torch.empty(need, dtype=torch.uint8, pin_memory=True)
        got = C.c_int64(0)
        t_host, t_dev, t_kern = [], [], []
        for it in range(a.iters + 2):
            torch.cuda.synchronize()
            t = time.perf_counter()
            h_hits.copy_(dh); h_rec.copy_(dr); torch.cuda.synchronize()
            host_fn(h_hits.data_ptr(), h_rec.data_ptr(), meta.ctypes.data, n, B.LAP_ANY, 0xFF, buf.ctypes.data, need)
            t_host.append(time.perf_counter() - t)
            t = time.perf_counter()
            B.check(L.btbb_b200_capture_records_dev(ctx.h, fmt, dh.data_ptr(), dr.data_ptr(), dm.data_ptr(), n, B.LAP_ANY, 0xFF,
                                                    d_out.data_ptr(), need, C.byref(got), None))
            t_kern.append(time.perf_counter() - t)
            h_out.copy_(d_out); torch.cuda.synchronize()
            t_dev.append(time.perf_counter() - t)
        assert got.value == need and h_out.numpy().tobytes() == buf.tobytes()
        out[name] = {"file_bytes": int(need), "record_bytes_d2h": int(n * (16 + 372)),
                     "host_route_ms": 1e3 * float(np.median(t_host[2:])), "device_route_ms": 1e3 * float(np.median(t_dev[2:])),
                     "device_kernels_ms": 1e3 * float(np.median(t_kern[2:])), "identical": True}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
