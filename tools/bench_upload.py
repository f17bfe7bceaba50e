"""Time beatgpu_upload_gflib: a float64 GF library in pageable host memory -> float32 rows in HBM (row f2).
Usage: python tools/bench_upload.py [GB]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beat_b200.lib import F32, Context  # noqa: E402


def main():
    gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    nt, npatch, ndur, ns = 8, 100, 17, 120
    nst = max(1, int(gb * 2**30 / (nt * npatch * ndur * ns * 8)))
    G = np.empty((nt, npatch, ndur, nst, ns))
    G[...] = np.arange(ns)
    ctx = Context(0)
    wid = ctx.add_wavemap(nt, ns, "multilinear", None, np.zeros(nt, np.int32), np.full(nt, ns, np.int32))
    for label in ("pageable source, pinned double buffering", "same (2nd run)"):
        t0 = time.perf_counter()
        ctx.upload_gflib(wid, 0, G, F32, 0.5, 0.25, -5.0, 0.5)
        dt = time.perf_counter() - t0
        print("%s: %.2f GB float64 -> float32 HBM rows in %.2f s = %.1f GB/s of source" % (label, G.nbytes / 1e9, dt, G.nbytes / 1e9 / dt))
    os.environ["BEATGPU_UPLOAD_DIRECT"] = "1"
    t0 = time.perf_counter()
    ctx.upload_gflib(wid, 0, G, F32, 0.5, 0.25, -5.0, 0.5)
    dt = time.perf_counter() - t0
    print("pageable source, plain cudaMemcpyAsync per chunk (previous behaviour): %.2f s = %.1f GB/s" % (dt, G.nbytes / 1e9 / dt))
    del os.environ["BEATGPU_UPLOAD_DIRECT"]
    ctx.pin(G)
    t0 = time.perf_counter()
    ctx.upload_gflib(wid, 0, G, F32, 0.5, 0.25, -5.0, 0.5)
    dt = time.perf_counter() - t0
    print("page-locked source: %.2f s = %.1f GB/s" % (dt, G.nbytes / 1e9 / dt))
    ctx.close()


if __name__ == "__main__":
    main()
