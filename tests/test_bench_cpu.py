"""CPU tests of bench.py's host side: the shared-memory host library of the CPU arm equals the plain recipe, and
`--impl reference` prints one JSON line that honours the contract (same metric / unit / config as the GPU arm, `impl`,
`cpu_baseline`, `e2e` with zero copy bytes) -- on the --quick shapes, so it runs in seconds without a GPU."""
import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_shared_host_library_equals_the_recipe(monkeypatch):
    bench = _bench()
    from beat_b200 import synthetic
    small = dict(nt=4, subfaults=((5, 8, 2.0),), ns=48, ndur=5, nst=24)
    monkeypatch.setattr(bench, "c3_args", lambda quick: dict(small))
    monkeypatch.setenv("BENCH_CPU_NDUR", "5")                     # = all duration nodes: the shared-memory path
    monkeypatch.setenv("BENCH_MAX_CORES", "3")
    args = types.SimpleNamespace(quick=False, interpolation="multilinear")
    prob = bench.build_cpu_problem(args)
    assert prob["host_library"].startswith("all 5 duration nodes")
    ref = synthetic.make_problem(interpolation="multilinear", seed=1234, **small)
    for v in prob["slip_vars"]:
        assert np.array_equal(prob["wavemaps"][0]["G"][v], ref["wavemaps"][0]["G"][v])
    assert np.array_equal(prob["wavemaps"][0]["data"], ref["wavemaps"][0]["data"])
    # fewer nodes when asked to (or when the box lacks the memory): same shapes otherwise, and the text says so
    monkeypatch.setenv("BENCH_CPU_NDUR", "2")
    prob2 = bench.build_cpu_problem(args)
    assert prob2["wavemaps"][0]["G"]["uparr"].shape == (4, 40, 2, 24, 48) and "2 of the 5" in prob2["host_library"]
    # the oracle evaluates on the shared library like on a private one
    from oracle import ffi_oracle as O
    q = synthetic.draw_chains(prob, 1, seed=2)[0]
    np.testing.assert_array_equal(O.ffi_seismic_eval(prob, synthetic.split_point(prob, q), impl="port"),
                                  O.ffi_seismic_eval(ref, synthetic.split_point(ref, q), impl="port"))


def test_reference_arm_line_contract():
    env = dict(os.environ, BENCH_MAX_CORES="2", PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--quick", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1                                        # ONE JSON line on stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["metric"] == "forward+loglike evals/sec (FFI seismic 200-patch)" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["gpu_launches"] == 0
    assert "workload" in line["config"] and line["n_gpus"] == 1 and line["steps"] == 2
