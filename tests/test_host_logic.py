"""Host-side logic of the drop-in modules that needs no GPU: parameter containers, bin construction,
stat_tools, simulation grid, metric mapping -- checked against fixtures produced by the reference
and (when the reference is present) against the live reference objects."""
import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN, golden


def test_matern_params_api_matches_reference_fixture():
    import model
    g = golden("params_api")
    p2, p1 = model.MaternParams(2), model.MaternParams(1)
    assert list(p2.get_names()) == list(g["names2"]) and list(p1.get_names()) == list(g["names1"])
    np.testing.assert_array_equal(p2.get_values().astype(float), g["values2"])
    np.testing.assert_array_equal(p1.get_values().astype(float), g["values1"])
    np.testing.assert_array_equal(np.array([list(b) for b in p2.get_bounds()], float), g["bounds2"])
    for name in ("sigma", "nu", "rho"):
        np.testing.assert_array_equal(getattr(p2, name).values, g[name + "2"])
    np.testing.assert_array_equal(p1.rho.values, g["rho1"])
    p2.set_values(g["set_vals"])
    np.testing.assert_array_equal(p2.nu.values, g["nu_after"])
    np.testing.assert_array_equal(p2.rho.values, g["rho_after"])
    np.testing.assert_array_equal(p2.get_values().astype(float), g["set_vals"])
    assert p2.n_params == 11 and p1.n_params == 4
    with pytest.raises(ValueError):
        p2.set_values(np.ones(4))
    with pytest.raises(AttributeError):
        p2.set_bounds(foo=(0, 1))
    p2.set_bounds(len_scale=(1e-8, 1.5))
    assert p2.to_dataframe().loc[5, "bounds"] == (1e-8, 1.5)
    np.testing.assert_array_equal(p2.reset_values().get_values().astype(float), g["values2"])


def test_params_live_reference(ref):
    import model
    for n in (1, 2):
        a, b = model.MaternParams(n), ref.model.MaternParams(n)
        pd.testing.assert_frame_equal(a.to_dataframe(), b.to_dataframe(), check_dtype=False)
        for name in ("sigma", "nu", "len_scale", "nugget", "rho"):
            np.testing.assert_array_equal(getattr(a, name).values, getattr(b, name).values)
            assert getattr(a, name).get_names() == getattr(b, name).get_names()


def test_bins_from_extrema_matches_reference_fixture():
    import fields
    g = golden("variogram_haversine_semivariogram")
    for key in ("00", "01", "11"):
        c = g["center" + key]
        centers, edges = fields._bins_from_extrema(c[0], c[-1], int(g["n_bins"]))
        np.testing.assert_array_equal(centers, c)
        np.testing.assert_array_equal(edges, g["edges" + key])
        assert edges[0] == 0 and len(edges) == len(centers) + 1


def test_construct_variogram_bins_signature():
    import fields
    d = np.array([0.0, 2.0, 5.0, 9.0, 4.0])
    centers, edges = fields._construct_variogram_bins(pd.DataFrame({"distance": d, "variogram": d}), 4)
    np.testing.assert_allclose(centers, np.linspace(2, 9, 4))
    assert edges[0] == 0 and len(edges) == 5


def test_stat_tools_match_reference_fixture():
    import stat_tools
    g = golden("stat_tools")
    np.testing.assert_allclose(stat_tools.simple_linear_regression(g["x"].copy()), g["slr"], rtol=1e-12, equal_nan=True)
    z, slope = stat_tools.detrend(g["x"].copy())
    np.testing.assert_allclose(z, g["detrended"], rtol=1e-10, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(slope, g["slope"], rtol=1e-12)
    assert abs(stat_tools.compute_xcor_1d(g["x"], g["y"], lag=0) - g["xcor0"]) < 1e-14
    assert abs(stat_tools.compute_xcor_1d(g["x"], g["y"], lag=2) - g["xcor2"]) < 1e-14
    assert np.isnan(stat_tools.compute_xcor_1d(g["x"], g["y"], lag=2, tau=100))
    np.testing.assert_allclose(stat_tools.compute_xcor_nd(g["Z1"], g["Z2"], lag=1, tau=10), g["xcor_nd"], rtol=1e-13,
                               equal_nan=True)
    np.testing.assert_array_equal(stat_tools.get_count(g["Z1"]), g["count"])
    allnan = np.full(5, np.nan)
    assert stat_tools.detrend(allnan)[0] is allnan and np.isnan(stat_tools.detrend(allnan)[1])


def test_cartesian_grid_layout():
    import sim
    g = golden("sim")
    grid = sim.CartesianGrid(xcount=12, ycount=12)
    np.testing.assert_array_equal(grid.coords.values, g["coords"])
    assert list(grid.coords.columns) == ["x", "y"] and grid.count == 144


def test_metric_mapping_and_errors():
    from cokrig_b200 import ops, METRIC_EUCLID, METRIC_HAVERSINE
    assert ops.metric_id("km", True) == METRIC_HAVERSINE
    assert ops.metric_id(None, False) == METRIC_EUCLID
    with pytest.raises(NotImplementedError):
        ops.metric_id("km", False)
    with pytest.raises(ValueError):
        ops._params(np.ones(5), 2)
    with pytest.raises(NotImplementedError):
        ops._params(np.ones(5), 3)
    assert ops.padded_ld(1) == 16 and ops.padded_ld(17) == 32 and ops.padded_ld(40000) == 40000


def test_predictor_process_count_mismatch():
    import fields, joint_prediction, model, point_prediction
    mf = fields.MultiField.from_arrays([np.zeros((3, 2))], [np.zeros(3)])
    for mod in (joint_prediction, point_prediction):
        with pytest.raises(ValueError, match="Number of theoretical processes"):
            mod.Predictor(model.MultivariateMatern(n_procs=2), mf)
    vc = fields.VarioConfig(10, 5, n_procs=1)
    est = fields.EmpiricalVariogram(pd.DataFrame(), vc, np.nan, [np.nan])
    with pytest.raises(ValueError, match="Number of theoretical processes"):
        model.MultivariateMatern(n_procs=2).fit(est)


def test_multifield_from_arrays_shapes():
    import fields
    mf = fields.MultiField.from_arrays([np.zeros((4, 2)), np.ones((3, 2))], [np.arange(4.0), np.arange(3.0)])
    assert mf.n_procs == 2 and mf.n_data == 7 and mf.fields[1].size == 3
    assert mf.fields[0].coords_main is mf.fields[0].coords
    assert np.isnan(mf.timestamp)


def test_wls_cost_helpers():
    import model
    y, f, c = np.array([1.0, 2.0, 3.0]), np.array([1.5, 0.0, 2.0]), np.array([10.0, 20.0, 30.0])
    assert model._wls(y[[0, 2]], f[[0, 2]], c[[0, 2]]) == pytest.approx(10 * (0.5 / 1.5) ** 2 + 30 * 0.25)
    assert model.MultivariateMatern._weighted_least_squares(y, f, c) == pytest.approx(
        10 * (0.5 / 1.5) ** 2 + 20 * 4.0 + 30 * 0.25)


# ------------------------------------------------------------------------------------------------ drop-in boundary
_REPLACED_PRIVATE = set()  # nothing is left out: the per-target helpers of point_prediction.Predictor exist as device-backed methods


def _resolve(key: str):
    """'mod.func' or 'mod.Class.method' in the drop-in modules (inherited methods count), or None."""
    import importlib
    parts = key.split(".")
    obj = importlib.import_module(parts[0])
    for name in parts[1:]:
        obj = getattr(obj, name, None)
        if obj is None:
            return None
    return obj if callable(obj) else None


def test_every_reference_callable_exists_with_the_same_signature():
    """Drop-in boundary: every function and method the reference defines in its six modules exists here with
    the same parameter names, kinds and defaults (fixture dumped from the unmodified reference by
    tests/golden/make_signatures.py)."""
    import inspect
    import json
    import os
    from conftest import GOLDEN
    ref_sigs = json.load(open(os.path.join(GOLDEN, "signatures.json")))
    assert len(ref_sigs) > 100
    missing = sorted(k for k in ref_sigs if _resolve(k) is None and k not in _REPLACED_PRIVATE)
    assert not missing, missing
    assert all(k.split(".")[-1].startswith("_") for k in _REPLACED_PRIVATE)
    for key, want in ref_sigs.items():
        fn = _resolve(key)
        if fn is None:
            continue
        got = [[p.name, p.kind.name, None if p.default is inspect._empty else repr(p.default)]
               for p in inspect.signature(fn).parameters.values()]
        assert got == want, key


def test_signature_fixture_is_current(ref):
    """The committed fixture equals what the reference on disk defines (build container only)."""
    import importlib.util
    import json
    import os
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_signatures", os.path.join(GOLDEN, "make_signatures.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    assert mk.collect(ref) == json.load(open(os.path.join(GOLDEN, "signatures.json")))


def test_bench_reference_arm_line_contract():
    """bench.py --impl reference prints ONE JSON line with the contract keys, uses every host core even when the
    launcher exported OMP_NUM_THREADS=1 (torchrun does), and never touches the GPU library."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--cpu-n", "300", "--cpu-m", "100", "--cpu-asm-n", "400"], capture_output=True, text=True,
                       env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] == d["e2e"]["value"]
    assert cb["value_without_verify"] > cb["value"] and set(cb["scaled_phases_s"]) == {"assemble_s", "verify_s", "factor_s", "solve_s"}
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the BLAS / LAPACK thread pools really are resized (torchrun exports OMP_NUM_THREADS=1 before Python starts)
    probe = ("import sys; sys.path.insert(0, %r); import bench, json, threadpoolctl; n = bench._all_host_threads(); "
             "print(json.dumps([n] + [p['num_threads'] for p in threadpoolctl.threadpool_info() if p['user_api'] == 'blas']))" % root)
    rp = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, env=env, timeout=120)
    assert rp.returncode == 0, rp.stderr[-2000:]
    counts = json.loads(rp.stdout.strip().splitlines()[-1])
    assert counts[0] == os.cpu_count() and len(counts) >= 2 and all(c == counts[0] for c in counts[1:]), counts
    env["RANK"] = "1"  # the other ranks exit 0 without work
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                        text=True, env=env, timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_preprocessing_and_back_transform_without_xarray():
    """the reference's preprocessing chain (src/fields.py:283-375) and back-transform
    (src/joint_prediction.py:155-205) on plain arrays, against the fixture tests/golden/make_golden_r2.py builds from the
    reference's numeric steps (its own simple_linear_regression + the same sklearn / numpy calls)."""
    import sys
    import pandas as pd
    sys.path.insert(0, GOLDEN)
    from make_golden_r2 import inputs_preprocess
    import fields
    import joint_prediction
    g = golden("preprocess")
    cube, lat, lon, ti = inputs_preprocess()
    np.testing.assert_allclose(fields.fit_linear_trend_array(cube), g["trend"], rtol=1e-13, equal_nan=True)
    pre = fields.preprocess_arrays(cube, lat, lon, ti)
    fr = pre["frame"].dropna(subset=["value"])
    np.testing.assert_array_equal(fr["lon"].values, g["lon"])  # (lon, lat) row order of a (lon, lat, time) dataset
    np.testing.assert_array_equal(fr["lat"].values, g["lat"])
    np.testing.assert_allclose(fr["value"].values, g["standardised"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(fr["spatial_trend"].values, g["spatial_trend"], rtol=1e-12)
    a = pre["attrs"]
    np.testing.assert_allclose(a["covariate_means"], g["covariate_means"], rtol=1e-14)
    np.testing.assert_allclose(a["covariate_scales"], g["covariate_scales"], rtol=1e-14)
    assert abs(a["spatial_mean"] - g["spatial_mean"]) < 1e-14 and abs(a["scale_fact"] / g["scale_fact"] - 1) < 1e-13
    assert abs(a["temporal_trend"] - g["trend"][ti]) < 1e-13
    # Field / MultiField built from the cube, then the back-transform of standardised predictions
    mf = fields.MultiField.from_cubes([cube, cube], [lat, lat], [lon, lon], [ti, ti])
    f = mf.fields[0]
    assert mf.n_procs == 2 and f.size == len(g["lon"]) and f.coords.shape == (f.size, 2)
    np.testing.assert_array_equal(f.coords[:, 0], g["lat"])
    np.testing.assert_allclose(f.values, g["standardised"], rtol=1e-12, atol=1e-13)
    assert abs(np.mean(f.values)) < 1e-12 and abs(np.std(f.values) - 1) < 1e-12
    P = object.__new__(joint_prediction.Predictor)
    P.mf, P.i, P.covariates = mf, 0, None
    back = P.postprocess_frame(pd.DataFrame({"lat": g["p_lat"], "lon": g["p_lon"], "pred": g["p_pred"], "pred_err": g["p_err"]}))
    np.testing.assert_allclose(back["pred"].values, g["back_pred"], rtol=1e-12)
    np.testing.assert_allclose(back["pred_err"].values, g["back_err"], rtol=1e-14)
    # round trip: back-transforming the standardised data at the data locations returns the original slice
    rt = P.postprocess_frame(pd.DataFrame({"lat": f.coords[:, 0], "lon": f.coords[:, 1], "pred": f.values, "pred_err": 0 * f.values}))
    orig = pd.DataFrame({"lon": np.repeat(lon, len(lat)), "lat": np.tile(lat, len(lon)), "v": cube[ti].T.ravel()}).dropna()
    np.testing.assert_allclose(rt["pred"].values, orig["v"].values, rtol=1e-12)
    # a covariate given as an array instead of lon / lat
    elev = np.add.outer(lat, -0.5 * lon)
    pre2 = fields.preprocess_arrays(cube, lat, lon, ti, {"elev": elev})
    assert pre2["covariate_names"] == ["elev"] and pre2["attrs"]["spatial_model"].coef_.shape == (1,)
