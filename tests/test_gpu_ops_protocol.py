"""The reference's own call sequence for the FFI seismic likelihood -- SeismicDistributerComposite.get_formula,
/root/reference/beat/models/seismic.py:1253-1349 -- replayed against the product's Op classes on the GPU, with the Ops
driven through pytensor's Op protocol (make_node -> perform -> declared-type / infer_shape check; tests/_op_protocol.py):

    sweepers[index](1 / velocities_sf, nuc_dip_idx, nuc_strike_idx) + time[index]      (:1263-1272)
    tile / repeat of the station corrections                                            (:1283-1296)
    gfs[key].stack_all(targetidxs=, starttimes=, durations=, slips=, interpolation=)    (:1317-1330), summed over slip vars
    residuals = data - synthetics                                                       (:1332)
    multivariate_normal_chol(datasets, weights, hyperparams, residuals, hp_specific=)   (:1335-1341)

and compared with the fused evaluator (one call) and the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from _op_protocol import op_protocol  # noqa: E402
from beat_b200 import synthetic  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402


class _Cov(object):
    def __init__(self, log_pdet):
        self.slog_pdet, self.log_pdet = None, log_pdet


class _Dataset(object):
    """What multivariate_normal_chol reads of a dataset (distributions.py:119-126)."""

    def __init__(self, samples, typ, log_pdet):
        self.samples, self.typ, self.covariance = samples, typ, _Cov(log_pdet)


@pytest.mark.parametrize("station_corrections", [False, True])
def test_get_formula_call_sequence_on_the_ops(station_corrections):
    prob = synthetic.make_problem(nt=5, subfaults=((4, 6, 2.0), (3, 5, 2.5)), ns=36, ndur=5, interpolation="multilinear", seed=21,
                                  station_corrections=station_corrections)
    wm = prob["wavemaps"][0]
    q = synthetic.draw_chains(prob, 1, seed=22)[0]
    point = synthetic.split_point(prob, q)
    ref = O.ffi_seismic_eval(prob, point, impl="port")
    with op_protocol() as (ops, geometry):
        from beat_b200.engine import BatchedFFILogLike
        # composite set-up (seismic.py:1097-1111): one Sweeper Op per subfault
        sweepers = [ops.Sweeper(h, nd, nstr, "c") for nd, nstr, h in prob["subfaults"]]
        gfs = ops.SeismicGFLibrary({v: wm["G"][v] for v in prob["slip_vars"]}, wm["dur_min"], wm["dur_step"], wm["st_min"], wm["st_step"])
        # ---- get_formula
        npatches = prob["npatches"]
        starttimes0 = np.zeros(npatches)
        cum = np.cumsum([0] + [nd * nstr for nd, nstr, _ in prob["subfaults"]])
        for index, (nd, nstr, h) in enumerate(prob["subfaults"]):
            nuc_dip_idx, nuc_strike_idx = O.fault_locations2idxs(point["nucleation_dip"][index], point["nucleation_strike"][index], h, h)
            out = sweepers[index](1.0 / point["velocities"][cum[index]:cum[index + 1]], nuc_dip_idx, nuc_strike_idx)
            assert out.owner.op is sweepers[index] and out.type.ndim == 1              # a graph variable of the declared type
            starttimes_tmp = out.eval() + point["time"][index]
            starttimes0[cum[index]:cum[index + 1]] = starttimes_tmp
        n_t = wm["nt"]
        if station_corrections:
            starttimes = (np.tile(starttimes0, n_t) - np.repeat(point["time_shifts"][wm["station_idx"]], npatches)).reshape((n_t, npatches))
        else:
            starttimes = np.tile(starttimes0, n_t).reshape((n_t, npatches))
        targetidxs = np.atleast_2d(np.arange(n_t)).T
        synthetics = np.zeros((n_t, wm["ns"]))
        for var in prob["slip_vars"]:
            synthetics += gfs.stack_all(targetidxs=targetidxs, starttimes=starttimes, durations=point["durations"], slips=point[var],
                                        interpolation=wm["interpolation"], component=var)
        residuals = wm["data"] - synthetics
        datasets = [_Dataset(int(wm["nsamples"][t]), "any_P_0_Z", float(wm["slog_pdet"][t])) for t in range(n_t)]
        hyperparams = {"h_any_P_0_Z": point["hypers"][wm["hyper_idx"][0]]}
        logpts = ops.multivariate_normal_chol(datasets, list(wm["U"]), hyperparams, residuals, hp_specific=False)
        np.testing.assert_allclose(logpts, ref, rtol=1e-10)
        # ---- the same evaluation as ONE fused Op (what replaces the sub-graph inside get_formula)
        ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
        fused = ops.FFILogLike(ev)
        lp_var, like_var = fused(q)
        assert lp_var.type.ndim == 1 and like_var.type.ndim == 0
        np.testing.assert_allclose(lp_var.eval(), logpts, rtol=1e-12)
        np.testing.assert_allclose(like_var.eval(), ref.sum(), rtol=1e-10)
        ev.close()
