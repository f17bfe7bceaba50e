"""CPU test of the batched trace writer: files are byte-identical to what the reference's NumpyChain writes
one record at a time (restated here from beat/backend.py:765-782,797-845)."""
import json
from collections import OrderedDict

import numpy as np

from beat_b200 import backend as bk


def _reference_style_file(path, var_shapes, var_dtypes, flat_names, lpoints):
    """What NumpyChain.setup + record_buffer produce for one chain (sequential appends)."""
    names = list(var_shapes.keys())
    dt = np.dtype({"names": names, "formats": ["{}{}".format(var_shapes[n], var_dtypes[n]) for n in names]})
    with open(path, "wb") as fh:
        fh.write((json.dumps({"flat_names": flat_names, "var_shapes": var_shapes, "var_dtypes": var_dtypes}) + "\n").encode())
    data = np.zeros(1, dtype=dt)
    with open(path, "ab+") as fh:
        for lp in lpoints:
            for n, arr in zip(names, lp):
                data[n] = arr
            data.tofile(fh)


def test_batched_chains_byte_compatible(tmp_path):
    rng = np.random.default_rng(0)
    var_shapes = OrderedDict([("uparr", (6,)), ("time", (1,)), ("seis_like", (4,)), ("like", ())])
    n_chains, n_steps = 5, 7
    w = bk.BatchedNumpyChains(str(tmp_path / "stage_0"), var_shapes, n_chains, buffer_size=3)
    w.setup()
    steps = []
    for _ in range(n_steps):
        vals = {k: rng.standard_normal((n_chains,) + s) for k, s in var_shapes.items()}
        steps.append(vals)
        w.write(vals)
    w.flush()
    assert w.stored_samples == n_steps
    for c in range(n_chains):
        ref_path = str(tmp_path / ("ref-%d.bin" % c))
        lpoints = [[st[k][c] for k in var_shapes] for st in steps]
        _reference_style_file(ref_path, var_shapes, w.var_dtypes, w.flat_names, lpoints)
        assert open(ref_path, "rb").read() == open(w.filename(c), "rb").read()
        got = bk.get_values(w.filename(c), "seis_like")
        assert got.shape == (n_steps, 4)
        np.testing.assert_array_equal(got, np.array([st["seis_like"][c] for st in steps]))
        np.testing.assert_array_equal(bk.get_values(w.filename(c), "like", burn=n_steps - 1)[0], steps[-1]["like"][c])
    assert bk.create_flat_names("x", (2, 2)) == ["x__0_0", "x__0_1", "x__1_0", "x__1_1"] and bk.create_flat_names("like", ()) == ["like"]
