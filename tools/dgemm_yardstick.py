"""Yardstick only (not a product path): what cuBLAS delivers for f64 GEMMs of the shapes `dgemm_tile_kernel` runs, on this
box.  The numbers give the FP64 tensor-pipe denominator for the GEMM rows in profiles/README.md.
Usage: python tools/dgemm_yardstick.py > out.json"""
import json
import torch


def rate(m, n, k, batch=1, iters=20):
    dev = torch.device("cuda", 0)
    a = torch.randn((batch, m, k), dtype=torch.float64, device=dev)
    b = torch.randn((batch, k, n), dtype=torch.float64, device=dev)
    c = torch.empty((batch, m, n), dtype=torch.float64, device=dev)
    for _ in range(3):
        torch.bmm(a, b, out=c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.bmm(a, b, out=c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"m": m, "n": n, "k": k, "batch": batch, "ms": ms, "TFLOPs": 2.0 * m * n * k * batch / (ms * 1e-3) / 1e12}


def main():
    out = {"device": torch.cuda.get_device_name(0), "what": "torch.bmm (cuBLAS) f64, full (not triangular) products", "results": [
        rate(8192, 8192, 8192, iters=5),           # large square: the pipe's ceiling
        rate(2048, 2000, 2048, batch=96),          # C2 dense misfit: Z_t = U_t R_t for 96 datasets, 2000 chains
        rate(120, 4000, 120, batch=64),            # C3 dense misfit
        rate(500, 2000, 400),                      # C4 geodetic mu = G^T slips (np*nvar = 400)
        rate(500, 2000, 500),                      # C4 geodetic Z = U R
    ]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
