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


def test_writer_threads_and_packed_records_write_the_same_bytes(tmp_path):
    """n_io_threads > 0 (two step buffers, background appends) and the packed-record path (`slot` / `commit`,
    `write_records`) produce exactly the files of the synchronous dict-fed writer, whatever the buffer size."""
    import pytest
    rng = np.random.default_rng(4)
    var_shapes = OrderedDict([("uparr", (6,)), ("time", (1,)), ("seis_like", (4,)), ("like", ())])
    n_chains, n_steps = 37, 11
    steps = [{k: rng.standard_normal((n_chains,) + s) for k, s in var_shapes.items()} for _ in range(n_steps)]
    ref = bk.BatchedNumpyChains(str(tmp_path / "sync"), var_shapes, n_chains, buffer_size=100)
    ref.setup()
    for v in steps:
        ref.write(v)
    ref.flush()
    assert ref.record_width == 12
    for tag, threads, buf, packed in (("thr", 3, 4, False), ("thr1", 1, 1, False), ("packed", 2, 3, True), ("packed_sync", 0, 5, True)):
        w = bk.BatchedNumpyChains(str(tmp_path / tag), var_shapes, n_chains, buffer_size=buf, n_io_threads=threads)
        w.setup()
        for i, v in enumerate(steps):
            if packed:
                rec = np.concatenate([np.asarray(v[k]).reshape(n_chains, -1) for k in var_shapes], axis=1)
                if i % 2:
                    w.write_records(rec)
                else:
                    w.slot()[...] = rec
                    w.commit()
            else:
                w.write(v)
        w.close()
        assert w.stored_samples == n_steps
        for c in range(n_chains):
            assert open(w.filename(c), "rb").read() == open(ref.filename(c), "rb").read(), (tag, c)
    mixed = bk.BatchedNumpyChains(str(tmp_path / "mixed"), OrderedDict([("x", (2,)), ("n", ())]), 3, var_dtypes={"n": "int32"})
    assert mixed.record_width is None
    with pytest.raises(TypeError):
        mixed.slot()
    # a writer thread's failure surfaces in the sampler's thread at the next flush / close
    w = bk.BatchedNumpyChains(str(tmp_path / "gone"), var_shapes, n_chains, buffer_size=2, n_io_threads=2)
    w.setup()
    w.dir_path = str(tmp_path / "does" / "not" / "exist")
    w.write(steps[0]); w.write(steps[1])
    with pytest.raises(OSError):
        w.close()


def _golden_trace():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "trace_golden.npz"))


def test_files_byte_identical_to_the_references_numpychain(tmp_path):
    """Fixture written by the REFERENCE'S OWN NumpyChain (beat/backend.py:651-897: setup -> write x 7 with a buffer of 3
    -> record_buffer; tests/golden/make_trace_golden.py): the batched writer produces the same bytes, file by file, and
    at generation time the reference's reader read our files back identically."""
    g = _golden_trace()
    assert bool(g["reference_reader_reads_our_files"])
    names = [str(n) for n in g["varnames"]]
    var_shapes = OrderedDict((n, tuple(int(x) for x in str(s).strip("()").split(",") if x.strip())) for n, s in zip(names, g["shapes"]))
    n_chains, n_steps = int(g["n_chains"]), int(g["n_steps"])
    for buffer_size in (int(g["buffer_size"]), 1, 100):                  # flush granularity must not matter
        w = bk.BatchedNumpyChains(str(tmp_path / ("b%d" % buffer_size)), var_shapes, n_chains, buffer_size=buffer_size)
        w.setup()
        for i in range(n_steps):
            w.write({n: g["step%d_%s" % (i, n)] for n in names})
        w.flush()
        for c in range(n_chains):
            ours = open(w.filename(c), "rb").read()
            assert ours == g["file_chain%d" % c].tobytes(), (buffer_size, c)
            assert ours.startswith(g["header"].tobytes())
    # and the reader half: the reference-written bytes parse into the values that were written
    for c in range(n_chains):
        path = str(tmp_path / ("ref-chain-%d.bin" % c))
        open(path, "wb").write(g["file_chain%d" % c].tobytes())
        for n in names:
            want = np.array([g["step%d_%s" % (i, n)][c] for i in range(n_steps)]).reshape((n_steps,) + var_shapes[n])
            np.testing.assert_array_equal(bk.get_values(path, n), want)


def test_smc_driver_streams_every_step_into_chain_files(tmp_path):
    """The lock-step SMC driver with the trace writer on its ``on_step`` hook: after the run every chain file of every
    stage holds n_steps records whose last one is that chain's end point (what the reference's stage directories hold,
    beat/backend.py:985-1156 / sampler/base.py:364,390)."""
    import torch
    from beat_b200 import sampler as S
    n, n_chains, n_steps = 3, 24, 6
    mu = torch.tensor([0.3, -0.2, 0.1], dtype=torch.float64)

    def evaluator(q):
        lp = -0.5 * ((q - mu) ** 2 / 0.05).sum(dim=1, keepdim=True)
        return lp, lp[:, 0]

    var_shapes = OrderedDict([("x", (n,)), ("seis_like", (1,)), ("like", ())])
    writers = {}

    def on_step(stage, step, q, logpts, like):
        if stage not in writers:
            w = bk.BatchedNumpyChains(str(tmp_path / ("stage_%d" % stage)), var_shapes, n_chains, buffer_size=4)
            w.setup()
            writers[stage] = w
        writers[stage].write({"x": q.cpu().numpy(), "seis_like": logpts.cpu().numpy(), "like": like.cpu().numpy()})

    res = S.smc_sample(evaluator, -np.ones(n), np.ones(n), n_chains, n_steps, seed=3, on_step=on_step)
    for w in writers.values():
        w.flush()
    assert len(writers) == res["n_stages"] and res["n_stages"] >= 2
    last = writers[max(writers)]
    for c in range(n_chains):
        x = bk.get_values(last.filename(c), "x")
        assert x.shape == (n_steps, n)
        np.testing.assert_array_equal(x[-1], res["population"][c])
        np.testing.assert_array_equal(bk.get_values(last.filename(c), "like")[-1], res["likelihoods"][c])
