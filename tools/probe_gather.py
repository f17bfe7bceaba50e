"""Measure the row-gather ceilings of this GPU (diagnostics; see csrc/probe.cuh) and print one JSON object.
Usage: python tools/probe_gather.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beat_b200.lib import Context  # noqa: E402

MODES = {0: "ldg128", 1: "tma_bulk+smem_read", 2: "tma_bulk_only", 3: "tma_batched16+smem_read", 4: "tma_batched16_only",
         5: "smem_local", 6: "dsmem_cluster2", 7: "dsmem_cluster4", 8: "dsmem_cluster8"}


def main():
    ctx = Context(0)
    out = {"device": ctx.device_info()[1], "results": []}
    # 480 B = one GF-library row of the FFI stack kernel; 8448 B = one GF-store window of the delay-and-sum kernel
    quick = "--quick" in sys.argv
    only = [int(x.split("=")[1]) for x in sys.argv if x.startswith("--mode=")]
    for row_bytes in ((480, 512, 960) if quick else (480, 512, 960, 2048, 8448)):
        for ws_mb in ((30,) if quick else (30, 512, 8192)):      # L2-resident (one stack-kernel chunk), > L2, >> L2
            for mode in (only or (0, 1, 2, 3, 4, 5, 6, 7, 8)):
                if mode >= 5 and ws_mb != 30:
                    continue                                   # shared-memory modes have no working set
                if mode in (3, 4) and row_bytes > 1024:
                    continue                                   # batched ring does not fit
                total = (24 << 30) if ws_mb < 1000 else (6 << 30)
                rows = max(8, total // (148 * 64 * row_bytes))
                gbs = max(ctx.probe_gather(mode, ws_mb << 20, row_bytes, rows, 3) for _ in range(2))
                out["results"].append({"row_bytes": row_bytes, "working_set_MB": ws_mb, "mode": MODES[mode], "GBps": round(gbs, 1)})
                print(out["results"][-1], file=sys.stderr)
    ctx.close()
    s = json.dumps(out, indent=1)
    print(s)
    outs = [x for x in sys.argv[1:] if not x.startswith("--")]
    if outs:
        open(outs[0], "w").write(s)


if __name__ == "__main__":
    main()
