"""Turn one `ncu --set full` capture of the stack kernels into profiles/stack_kernel_traffic.json -- the record bench.py
reads for `roofline.traffic`.  The record carries the sha256 of csrc/stack.cuh and the L2 blocking the capture ran with;
bench.py reports the traffic only when both match the running build, otherwise null.

Usage: python tools/capture_traffic.py <stack.ncu-rep> <bench line json of the same build> [<misfit.ncu-rep>] > profiles/stack_kernel_traffic.json
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beat_b200.build import kernel_hash  # noqa: E402


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(units, r))) for r in rows[2:]]


def val(row, key):
    unit, v = row[key]
    v = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(unit, 1.0)
    return v * scale


def main():
    rep, bench_json = sys.argv[1], sys.argv[2]
    line = json.loads(open(bench_json).read().strip().splitlines()[-1])
    stack = [r for r in raw_rows(rep) if "gf_stack_chunk_kernel" in r["Kernel Name"][1]][0]
    out = {
        "source": "%s (ncu --set full --clock-control none, one launch of bench.py --steps 2 --warmup 3, 1x B200)" % os.path.basename(rep),
        "chains": line["config"]["chains_per_gpu"], "store": line["config"]["gf_storage"], "interpolation": line["config"]["interpolation"],
        "stack_cuh_hash": kernel_hash(), "l2_blocking": line["roofline"]["l2_blocking"],
        "kernel": stack["Kernel Name"][1],
        "dram_bytes_read": val(stack, "dram__bytes_read.sum"), "dram_bytes_write": val(stack, "dram__bytes_write.sum"),
        "lts_sectors_read": val(stack, "lts__t_sectors_srcunit_tex_op_read.sum"),
        "lts_hit_rate_pct": val(stack, "lts__t_sector_hit_rate.pct"),
        "l1tex_throughput_pct": val(stack, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        "lts_throughput_pct": val(stack, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "kernel_ms_under_ncu": val(stack, "gpu__time_duration.sum"),
    }
    out["dram_bytes_per_launch"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    out["l2_to_sm_bytes"] = out["lts_sectors_read"] * 32
    if len(sys.argv) > 3:
        mis = [r for r in raw_rows(sys.argv[3]) if "misfit" in r["Kernel Name"][1] and "gf_stack" not in r["Kernel Name"][1]]
        if mis:
            out["misfit_kernel_dram_bytes"] = val(mis[0], "dram__bytes_read.sum") + val(mis[0], "dram__bytes_write.sum")
            out["dram_bytes_per_launch"] += out["misfit_kernel_dram_bytes"]
    out["note"] = ("L2-blocked gather: all chains stream through one (target, patch-chunk) working set while it is L2 resident, so "
                   "DRAM sees about one pass over the touched library instead of the bytes the chains request; the L2->SM "
                   "return path carries l2_to_sm_bytes per launch.")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
