"""
On-disk contract of the lev/*.py scripts: the
runner functions write the files the reference's plotting stage loads, with
the reference's names, shapes and dtype, and their content equals what the
unmodified reference functions produced on the same outcomes (fixtures).
"""
import os

import numpy as np
import pytest
import torch

import golden_io
from test_oracle_lev import oracle_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scripts():
    from rlmd_b200 import lev_scripts
    return lev_scripts


def _load_all(path):
    return {f[:-4]: np.load(os.path.join(path, f)) for f in sorted(os.listdir(path)) if f.endswith(".npy")}


def test_coin_flip_files(scripts, tmp_path, capsys):
    n, h = 20_000, 40
    out = scripts.coin_flip(str(tmp_path) + "/", investors=n, horizon=h, s3=(0.25, 0.75, 0.25), r3=(0.70, 0.90, 0.10),
                            ru=(0.2, 0.8, 0.2), rd=(0.2, 0.8, 0.2))
    files = _load_all(tmp_path)
    assert sorted(files) == ["coin_inv1_val", "coin_inv1_val_T", "coin_inv2_val", "coin_inv3_val", "coin_inv4_lev"]
    assert files["coin_inv1_val"].shape == (10, 13, h - 1) and files["coin_inv1_val_T"].shape == (10, n)
    assert files["coin_inv2_val"].shape == (1, 1, 26, h - 1)
    from rlmd_b200.lev_exp import param_range                      # the reference's grid quirks (App. A)
    assert files["coin_inv3_val"].shape == (len(param_range(0.70, 0.90, 0.10)), len(param_range(0.25, 0.75, 0.25)),
                                            26, h - 1)
    assert files["coin_inv4_lev"].shape[0] == 3 and files["coin_inv4_lev"].shape[3] == 4
    for k, a in files.items():
        assert a.dtype == np.float32, k
        assert np.array_equal(a, out[k], equal_nan=True)
    # the leverage row of inv1 is the grid of the script (lev/coin_flip.py:72)
    assert np.allclose(files["coin_inv1_val"][:, 12, 0], np.arange(1, 11) / 10, atol=1e-6)
    text = capsys.readouterr().out
    assert "lev 5%:" in text and "TOTAL TIME" in text


def test_dice_roll_and_sh_files(scripts, tmp_path):
    n, h = 10_000, 24
    scripts.dice_roll(str(tmp_path) + "/", investors=n, horizon=h, s3=(0.5, 0.5, 0.1), r3=(0.45, 0.55, 0.05))
    scripts.dice_roll_sh(str(tmp_path) + "/", investors=n, horizon=h)
    files = _load_all(tmp_path)
    assert sorted(files) == ["dice_inv1_val", "dice_inv1_val_T", "dice_inv2_val", "dice_inv3_val",
                             "dice_sh_inv1_val", "dice_sh_inv1_val_T"]
    assert files["dice_inv1_val"].shape == (10, 13, h - 1) and files["dice_inv1_val_T"].shape == (10, n)
    assert files["dice_inv3_val"].shape[2:] == (26, h - 1)
    assert files["dice_sh_inv1_val"].shape == (10, 13, h - 1)
    assert all(a.dtype == np.float32 for a in files.values())


def test_gbm_files(scripts, tmp_path):
    n, h = 10_000, 33
    scripts.gbm(str(tmp_path) + "/", investors=n, horizon=h)
    files = _load_all(tmp_path)
    assert sorted(files) == ["gbm_op_inv1_val", "gbm_op_inv1_val_T", "gbm_snp_inv1_val", "gbm_snp_inv1_val_T"]
    for name in ("gbm_op", "gbm_snp"):
        assert files[name + "_inv1_val"].shape == (10, 13, h - 1)
        assert files[name + "_inv1_val_T"].shape == (10, n)
    # time-average growth of the unlevered-ish paths sits near mu - sigma^2/2 scaled by the leverage
    g = np.log(files["gbm_snp_inv1_val_T"].astype(np.float64) / 100.0).mean(axis=1) / h
    lev = files["gbm_snp_inv1_val"][:, 12, 0].astype(np.float64)
    mu, sg = 0.0540025395205692, 0.1897916175617430
    assert np.allclose(g, lev * (mu - sg ** 2 / 2), atol=4 * sg * lev.max() / np.sqrt(n * h))


@pytest.mark.parametrize("name", ["coin_testscale", "dice_testscale", "dicesh_testscale"])
def test_injected_outcomes_reproduce_the_reference_files(scripts, tmp_path, name):
    """With the fixture's outcomes injected, the saved inv1 arrays equal the reference's."""
    case = golden_io.lev_case(name)
    oc, f, lev, _ = oracle_inputs(case)
    gold = golden_io.load("lev_" + name, oc)
    codes = torch.from_numpy(oc).cuda()
    kw = dict(investors=case["n"], horizon=case["h"], value_0=case["v0"], l0=case["grid"], l1=case["grid"])
    # the fixtures fix `top` themselves; the scripts derive it from INVESTORS - only run matching cases
    top = int(case["n"] * 1e-4) if case["n"] * 1e-4 > 1 else 1
    if top != case["top"]:
        pytest.skip("fixture uses a top-K the scripts' rule does not produce")
    if case["kind"] == "coin":
        out = scripts.coin_flip(None, outcomes=codes, galaxy=False, up_r=case["up_r"], down_r=case["down_r"],
                                s2=(0.1, 0.1, 0.1), r2=(0.0, 0.0, 0.1), s3=(0.5, 0.5, 0.1), r3=(0.7, 0.7, 0.1), **kw)
        key = "coin"
    elif case["kind"] == "dice":
        out = scripts.dice_roll(None, outcomes=codes, s3=(0.5, 0.5, 0.1), r3=(0.7, 0.7, 0.1), **kw)
        key = "dice"
    else:
        out = scripts.dice_roll_sh(None, outcomes=codes, **kw)
        key = "dice_sh"
    got_T = out[key + "_inv1_val_T"]
    assert np.array_equal(got_T.view(np.uint32), gold["data_T"].view(np.uint32))
    cols = golden_io.kept_columns(case)
    got = out[key + "_inv1_val"][:, :, cols]
    assert np.array_equal(got[:, 9:13], gold["data"][:, 9:13])          # medians and the leverage row: exact
    assert np.allclose(got[:, :9], gold["data"][:, :9], rtol=3e-5, equal_nan=True)


def test_inv1_plot_inputs_match_numpy_on_reference_arrays(scripts):
    """The derived arrays are plain NumPy on the saved files; checked on a fixture's data."""
    case = golden_io.lev_case("coin_top7")
    oc, f, lev, _ = oracle_inputs(case)
    gold = golden_io.load("lev_" + case["name"], oc)
    d = scripts.inv1_plot_inputs(gold["data"], gold["data_T"], 1e30)
    assert d["log_vals_95"].shape == (gold["data_T"].shape[0],)
    want = np.log10(np.percentile(gold["data_T"], 5, method="median_unbiased", axis=1))
    assert np.allclose(d["log_vals_95"], want)
    assert d["mean_adj_v"].shape == gold["data"][:, 2].shape
    assert np.allclose(d["nor_mad_up"], np.log10(np.minimum(1e30, gold["data"][:, 0, -1] + gold["data"][:, 3, -1])))
