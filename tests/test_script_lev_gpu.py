"""
The reference's integration harness (tests/test_script_lev.py:196-494: every `lev.lev_exp`
function at N = 1e4, H = 3e2, seed 420, saved / reloaded / handed to the plot functions) run
against the engine through the injection INTEGRATION.md prescribes:

    sys.modules["lev.lev_exp"] = rlmd_b200.lev_exp

The harness body below restates the reference's sequence of calls (same arguments, same
order of random draws from torch's CPU generator, same file names); the module it imports
as `lev.lev_exp` is the engine's.  What it saves is compared with what the UNMODIFIED
reference harness saved (tests/golden/script_lev.npz, written by
tests/golden/gen_golden_script_lev.py, which executes the reference file where it lies).
matplotlib is absent on both sides: `plotting.plots_multiverse` is a stub that records calls.
"""
import contextlib
import hashlib
import io
import os
import sys
import types

import numpy as np
import pytest
import torch as T

import golden_io
from test_oracle_lev import assert_stats_close

pytestmark = pytest.mark.gpu


def harness(path_results, path_figs):
    """The calls of tests/test_script_lev.py:196-494, in order (VRAM = False: CPU tensors, device cpu)."""
    from torch.distributions.bernoulli import Bernoulli
    from torch.distributions.categorical import Categorical
    from torch.distributions.normal import Normal

    import plotting.plots_multiverse as plots
    from lev.lev_exp import (coin_big_brain_lev, coin_fixed_final_lev, coin_galaxy_brain_lev, coin_smart_lev,
                             dice_big_brain_lev, dice_fixed_final_lev, dice_sh_fixed_final_lev, dice_sh_smart_lev,
                             dice_smart_lev, gbm_fixed_final_lev, gbm_smart_lev)

    device = T.device("cpu")
    T.manual_seed(420)
    INVESTORS = T.tensor(int(1e4), dtype=T.int32, device=device)
    HORIZON = T.tensor(int(3e2), dtype=T.int32, device=device)
    VALUE_0 = T.tensor(1e2, device=device)
    TOP = 1
    ASYM_LIM = T.tensor(1e-12, device=device)
    l0 = l1 = (0.50, 1.00, 0.10)
    s2, r2 = (0.10, 0.10, 0.10), (0.00, 0.00, 0.10)
    s3, r3 = (0.70, 0.80, 0.10), (0.70, 0.80, 0.10)

    def lev_factor(up_r, down_r):
        bigger = np.abs(down_r) if np.abs(up_r) >= np.abs(down_r) else -np.abs(up_r)
        f = T.tensor(1 / bigger, device=device)
        return f - ASYM_LIM if np.abs(up_r) > np.abs(down_r) else f + ASYM_LIM

    def save(name, t):
        np.save(path_results + name + ".npy", t.cpu().numpy())

    # COIN FLIP (:204-311)
    up, dn = 0.5, -0.4
    LEV_FACTOR = lev_factor(up, dn)
    outcomes = Bernoulli(0.5).sample(sample_shape=(INVESTORS, HORIZON)).to(device)
    coin_fixed_final_lev(device, outcomes, TOP, VALUE_0, up, dn, lev_low=l0[0], lev_high=l0[1], lev_incr=l0[2])
    d, dT = coin_smart_lev(device, outcomes, INVESTORS, HORIZON, TOP, VALUE_0, up, dn, lev_low=l1[0], lev_high=l1[1],
                           lev_incr=l1[2])
    save("coin_inv1_val", d)
    save("coin_inv1_val_T", dT)
    for tag, s, r in (("coin_inv2_val", s2, r2), ("coin_inv3_val", s3, r3)):
        save(tag, coin_big_brain_lev(device, outcomes, INVESTORS, HORIZON, TOP, VALUE_0, up, dn, LEV_FACTOR,
                                     stop_min=s[0], stop_max=s[1], stop_incr=s[2], roll_min=r[0], roll_max=r[1],
                                     roll_incr=r[2]))
    save("coin_inv4_lev", coin_galaxy_brain_lev(device, ru_min=0.50, ru_max=0.80, ru_incr=0.10, rd_min=0.50,
                                                rd_max=0.80, rd_incr=0.10, pu_min=0.25, pu_max=0.75, pu_incr=0.25))
    plots.plot_inv4(np.load(path_results + "coin_inv4_lev.npy"), path_figs + "coin_inv4")
    plots.plot_inv3(np.load(path_results + "coin_inv3_val.npy"), path_figs + "coin_inv3")
    plots.plot_inv2(np.load(path_results + "coin_inv2_val.npy"), 30, path_figs + "coin_inv2")
    plots.plot_inv1(np.load(path_results + "coin_inv1_val.npy"), np.load(path_results + "coin_inv1_val_T.npy"), 1e30,
                    path_figs + "coin_inv1")

    # DICE ROLL (:313-401)
    up, dn, mid = 0.5, -0.5, 0.05
    probs = T.tensor([1 / 6, 1 / 6, 1 - (1 / 6 + 1 / 6)], device=device)
    LEV_FACTOR = lev_factor(up, dn)
    outcomes = Categorical(probs).sample(sample_shape=(INVESTORS, HORIZON)).to(device)
    dice_fixed_final_lev(device, outcomes, TOP, VALUE_0, up, dn, mid, lev_low=l0[0], lev_high=l0[1], lev_incr=l0[2])
    d, dT = dice_smart_lev(device, outcomes, INVESTORS, HORIZON, TOP, VALUE_0, up, dn, mid, lev_low=l1[0],
                           lev_high=l1[1], lev_incr=l1[2])
    save("dice_inv1_val", d)
    save("dice_inv1_val_T", dT)
    for tag, s, r in (("dice_inv2_val", s2, r2), ("dice_inv3_val", s3, r3)):
        save(tag, dice_big_brain_lev(device, outcomes, INVESTORS, HORIZON, TOP, VALUE_0, up, dn, mid, LEV_FACTOR,
                                     stop_min=s[0], stop_max=s[1], stop_incr=s[2], roll_min=r[0], roll_max=r[1],
                                     roll_incr=r[2]))
    plots.plot_inv3(np.load(path_results + "dice_inv3_val.npy"), path_figs + "dice_inv3")
    plots.plot_inv2(np.load(path_results + "dice_inv2_val.npy"), 90, path_figs + "dice_inv2")
    plots.plot_inv1(np.load(path_results + "dice_inv1_val.npy"), np.load(path_results + "dice_inv1_val_T.npy"), 1e40,
                    path_figs + "dice_inv1")

    # DICE ROLL, SAFE HAVEN (:403-448)
    sh = (-1, 5, -1)
    outcomes = Categorical(T.tensor([1 / 6, 1 / 6, 1 - (1 / 6 + 1 / 6)], device=device)).sample(
        sample_shape=(INVESTORS, HORIZON)).to(device)
    dice_sh_fixed_final_lev(device, outcomes, TOP, VALUE_0, up, dn, mid, *sh, lev_low=l0[0], lev_high=l0[1],
                            lev_incr=l0[2])
    d, dT = dice_sh_smart_lev(device, outcomes, INVESTORS, HORIZON, TOP, VALUE_0, up, dn, mid, *sh, lev_low=l1[0],
                              lev_high=l1[1], lev_incr=l1[2])
    save("dice_sh_inv1_val", d)
    save("dice_sh_inv1_val_T", dT)
    plots.plot_inv1(np.load(path_results + "dice_sh_inv1_val.npy"), np.load(path_results + "dice_sh_inv1_val_T.npy"),
                    1e40, path_figs + "dice_sh_inv1")

    # GEOMETRIC BROWNIAN MOTION (:450-494)
    drift, vol, names = [0.05, 0.0540025395205692], [np.sqrt(0.2), 0.1897916175617430], ["gbm_op", "gbm_snp"]
    g0 = ([-1.0, 0.4], [1.0, 4.0], [0.2, 0.4])
    g1 = ([-1.0, 0.2], [1.0, 2.0], [0.2, 0.2])
    for x in range(2):
        LOG_MEAN = T.tensor(drift[x] - vol[x] ** 2 / 2, device=device)
        VOL = T.tensor(vol[x], device=device)
        outcomes = Normal(LOG_MEAN, VOL).sample(sample_shape=(INVESTORS, HORIZON)).to(device)
        gbm_fixed_final_lev(device, outcomes, TOP, VALUE_0, lev_low=g0[0][x], lev_high=g0[1][x], lev_incr=g0[2][x])
        d, dT = gbm_smart_lev(device, outcomes, INVESTORS, HORIZON, TOP, VALUE_0, lev_low=g1[0][x], lev_high=g1[1][x],
                              lev_incr=g1[2][x])
        save(names[x] + "_inv1_val", d)
        save(names[x] + "_inv1_val_T", dT)
        plots.plot_inv1(np.load(path_results + names[x] + "_inv1_val.npy"),
                        np.load(path_results + names[x] + "_inv1_val_T.npy"), 1e40, path_figs + names[x] + "_inv1")


def test_reference_harness_through_the_injected_module(tmp_path, monkeypatch):
    import rlmd_b200.lev_exp as engine_lev_exp

    gold = np.load(os.path.join(golden_io.GOLDEN_DIR, "script_lev.npz"), allow_pickle=False)
    # ---- the injection and the plot stub
    lev_pkg = types.ModuleType("lev")
    lev_pkg.lev_exp = engine_lev_exp
    calls = []
    plots = types.ModuleType("plotting.plots_multiverse")
    for name in ("plot_inv1", "plot_inv2", "plot_inv3", "plot_inv4"):
        setattr(plots, name, (lambda n: (lambda *a, **k: calls.append(f"{n}:{os.path.basename(a[-1])}")))(name))
    plot_pkg = types.ModuleType("plotting")
    plot_pkg.plots_multiverse = plots
    for k, v in (("lev", lev_pkg), ("lev.lev_exp", engine_lev_exp), ("plotting", plot_pkg),
                 ("plotting.plots_multiverse", plots)):
        monkeypatch.setitem(sys.modules, k, v)
    res, figs = str(tmp_path / "results") + os.sep, str(tmp_path / "figs") + os.sep
    os.makedirs(res)
    os.makedirs(figs)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        harness(res, figs)
    assert calls == [str(c) for c in gold["plot_calls"]]

    # ---- what the harness saved against what the reference's harness saved
    saved = {f[:-4]: np.load(os.path.join(res, f)) for f in os.listdir(res)}
    assert sorted(saved) == sorted(k for k in ("coin_inv1_val coin_inv1_val_T coin_inv2_val coin_inv3_val coin_inv4_lev "
                                               "dice_inv1_val dice_inv1_val_T dice_inv2_val dice_inv3_val "
                                               "dice_sh_inv1_val dice_sh_inv1_val_T gbm_op_inv1_val gbm_op_inv1_val_T "
                                               "gbm_snp_inv1_val gbm_snp_inv1_val_T").split())
    for a in saved.values():
        assert a.dtype == np.float32
    assert np.array_equal(saved["coin_inv4_lev"], gold["coin_inv4_lev"])
    for g in ("coin", "dice", "dice_sh"):
        dT = saved[g + "_inv1_val_T"]
        assert hashlib.sha256(np.ascontiguousarray(dT).tobytes()).hexdigest() == str(gold[g + "_inv1_val_T_sha256"]), g
        d, cols = saved[g + "_inv1_val"], gold[g + "_inv1_val_cols"]
        want = gold[g + "_inv1_val"]
        assert d.shape == (6, 13, 299)
        assert np.array_equal(d[:, 9:13][..., cols].view(np.uint32), want[:, 9:13].view(np.uint32))   # medians, leverage
        assert_stats_close(d[:, :9][..., cols], want[:, :9])
    for g in ("gbm_op", "gbm_snp"):
        dT, want_T = saved[g + "_inv1_val_T"][:, ::10], gold[g + "_inv1_val_T_sub"]
        ok = np.isfinite(want_T) & (want_T > 1e-30)
        assert (np.abs(dT[ok].astype(np.float64) - want_T[ok]) <= 1e-5 * want_T[ok]).all()
        d, cols = saved[g + "_inv1_val"], gold[g + "_inv1_val_cols"]
        assert d.shape == (10, 13, 299) and np.array_equal(d[:, 12][..., cols], gold[g + "_inv1_val"][:, 12])
        assert_stats_close(d[:, :12][..., cols], gold[g + "_inv1_val"][:, :12], rtol=2e-5)
    for g in ("coin", "dice"):
        for inv, shape in (("inv2", (1, 1, 26, 299)), ("inv3", (3, 3, 26, 299))):
            d, cols, want = saved[f"{g}_{inv}_val"], gold[f"{g}_{inv}_val_cols"], gold[f"{g}_{inv}_val"]
            assert d.shape == shape
            d = d[..., cols]
            assert np.array_equal(d[:, :, 24:26], want[:, :, 24:26])                      # stop-loss / retention rows
            for lo_ in (0, 12):    # wealth statistics, leverage statistics: medians exact, moments to fp32 noise
                assert np.array_equal(d[:, :, lo_ + 9:lo_ + 12].view(np.uint32),
                                      want[:, :, lo_ + 9:lo_ + 12].view(np.uint32)), (g, inv, lo_)
                assert_stats_close(d[:, :, lo_:lo_ + 9], want[:, :, lo_:lo_ + 9], rtol=3e-5, noise=2e-6)

    # ---- the printed report, line for line (the GBM lines to the printed precision's last digit)
    got, want = buf.getvalue().splitlines(), str(gold["text"]).splitlines()
    # the reference file also prints its input validator's lines first and a footer last
    first = next(i for i, l in enumerate(want) if l.startswith("       lev "))
    want = [l for l in want[first:] if not l.startswith(("TOTAL TIME", "-----", "All Leverage"))]
    assert len(got) == len(want)
    diff = [(a, b) for a, b in zip(got, want) if a != b]
    # the *_smart_lev / big-brain reports print fp32 moments of the engine's fp64 accumulation:
    # a line may differ in the last printed digit of a mean / mad / std; order statistics never
    def close(a, b):
        ta, tb = a.split(), b.split()
        if len(ta) != len(tb):
            return False
        for x, y in zip(ta, tb):
            if x == y:
                continue
            try:
                fx, fy = float(x), float(y)
            except ValueError:
                return False
            if "e" in y:      # d.dde+xx: one unit of the last mantissa digit
                unit = 10.0 ** (np.floor(np.log10(max(abs(fy), 1e-300))) - len(y.split("e")[0].lstrip("-").split(".")[1]))
            else:             # plain decimals
                unit = 10.0 ** (-len(y.split(".")[1])) if "." in y else 1.0
            if not abs(fx - fy) <= 1.1 * unit:
                return False
        return True
    assert all(close(a, b) for a, b in diff), diff[:5]
    assert len(diff) <= len(want) // 20, (len(diff), diff[:5])
