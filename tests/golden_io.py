"""
Case tables and input regeneration shared by tests/golden/gen_golden.py (which
runs the reference) and the parity tests (which do not).
"""
import hashlib
import math
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SH = (-1.0, 5.0, -1.0)

# grids are (low, high, incr) as the reference scripts pass them
LEV_CASES = [
    # the reference's own smoke-test scale (tests/test_script_lev.py:61-66,99-100)
    dict(name="coin_testscale", kind="coin", n=10000, h=300, top=1, v0=1e2, seed=420,
         up_r=0.5, down_r=-0.4, p=(0.5,), grid=(0.5, 1.0, 0.1), stride=13),
    dict(name="dice_testscale", kind="dice", n=10000, h=300, top=1, v0=1e2, seed=421,
         up_r=0.5, down_r=-0.5, mid_r=0.05, p=(1 / 6, 1 / 6), grid=(0.5, 1.0, 0.1), stride=13),
    dict(name="dicesh_testscale", kind="dice_sh", n=10000, h=300, top=1, v0=1e2, seed=422,
         up_r=0.5, down_r=-0.5, mid_r=0.05, sh=SH, p=(1 / 6, 1 / 6), grid=(0.5, 1.0, 0.1), stride=13),
    dict(name="gbm_testscale", kind="gbm", n=10000, h=300, top=1, v0=1e2, seed=423,
         mu=0.05, sigma=math.sqrt(0.2), grid=(-1.0, 1.0, 0.2), stride=13),
    # top-K > 1, ragged sizes (N not a multiple of 32/128, odd H)
    dict(name="coin_top7", kind="coin", n=2001, h=41, top=7, v0=1e2, seed=1,
         up_r=0.5, down_r=-0.4, p=(0.5,), grid=(0.1, 1.0, 0.1), stride=1),
    dict(name="dice_top5", kind="dice", n=1537, h=67, top=5, v0=1e2, seed=2,
         up_r=0.5, down_r=-0.5, mid_r=0.05, p=(1 / 6, 1 / 6), grid=(0.05, 1.0, 0.05), stride=1),
    dict(name="dicesh_top3", kind="dice_sh", n=1030, h=50, top=3, v0=1e2, seed=3,
         up_r=0.5, down_r=-0.5, mid_r=0.05, sh=SH, p=(1 / 6, 1 / 6), grid=(0.73, 1.0, 0.03), stride=1),
    dict(name="gbm_snp_top4", kind="gbm", n=1999, h=64, top=4, v0=1e2, seed=4,
         mu=0.0540025395205692, sigma=0.1897916175617430, grid=(0.2, 2.0, 0.2), stride=1),
    # sign flip of the grid (-down_r > up_r), lev/lev_exp.py:81,161
    dict(name="coin_flip_sign", kind="coin", n=777, h=33, top=2, v0=1e2, seed=5,
         up_r=0.3, down_r=-0.5, p=(0.6,), grid=(0.2, 0.8, 0.2), stride=1),
    # edge sizes: two steps only, a handful of investors
    dict(name="dice_tiny", kind="dice", n=5, h=2, top=1, v0=1e2, seed=6,
         up_r=0.5, down_r=-0.5, mid_r=0.05, p=(1 / 6, 1 / 6), grid=(0.5, 1.0, 0.1), stride=1),
    # long horizon: fp32 underflow towards denormals / zero at full leverage
    dict(name="coin_long", kind="coin", n=513, h=3000, top=1, v0=1e2, seed=7,
         up_r=0.5, down_r=-0.4, p=(0.5,), grid=(0.2, 1.0, 0.2), stride=250),
    # fp32 overflow to inf for part of the investors (torch.std_mean -> nan there)
    dict(name="coin_overflow", kind="coin", n=400, h=1000, top=3, v0=1e2, seed=9,
         up_r=0.5, down_r=-0.4, p=(0.7,), grid=(0.2, 1.0, 0.4), stride=37),
    # GBM underflow to zero in fp32
    dict(name="gbm_overflow", kind="gbm", n=300, h=4000, top=1, v0=1e2, seed=8,
         mu=0.05, sigma=math.sqrt(0.2), grid=(0.2, 2.0, 0.6), stride=500),
]


def draw_outcomes(case: dict) -> np.ndarray:
    """uint8 codes [N,H] (coin: 1 = up; dice: 0 up, 1 down, 2 mid) or fp32 normals."""
    rs = np.random.RandomState(case["seed"])
    n, h = case["n"], case["h"]
    if case["kind"] == "coin":
        return (rs.random_sample((n, h)) < case["p"][0]).astype(np.uint8)
    if case["kind"] in ("dice", "dice_sh"):
        u = rs.random_sample((n, h))
        p_up, p_dn = case["p"]
        return np.where(u < p_up, 0, np.where(u < p_up + p_dn, 1, 2)).astype(np.uint8)
    if case["kind"] == "gbm":
        mean = np.float32(case["mu"] - case["sigma"] ** 2 / 2)
        z = rs.standard_normal((n, h)).astype(np.float32)
        return (mean + np.float32(case["sigma"]) * z).astype(np.float32)
    raise ValueError(case["kind"])


def kept_columns(case: dict):
    """Time columns of data[G,13,H-1] that the fixture keeps."""
    hm1 = case["h"] - 1
    cols = list(range(0, hm1, case["stride"]))
    if cols[-1] != hm1 - 1:
        cols.append(hm1 - 1)
    return cols


def load(name: str, inputs: np.ndarray = None) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    out = {k: z[k] for k in z.files}
    if inputs is not None:
        got = hashlib.sha256(np.ascontiguousarray(inputs).tobytes()).hexdigest()
        assert got == str(out["input_sha256"]), "regenerated inputs differ from the fixture's"
    return out


def lev_case(name: str) -> dict:
    for c in LEV_CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)


# ---------------------------------------------------------------------- envs
# (fixture name, reference module, class, family, investor, n_gambles)
ENV_CASES = [
    ("coin_A1", "envs.coin_flip_envs", "Coin_InvA", "coin", "A", 1),
    ("coin_A3", "envs.coin_flip_envs", "Coin_InvA", "coin", "A", 3),
    ("coin_B1", "envs.coin_flip_envs", "Coin_InvB", "coin", "B", 1),
    ("coin_B2", "envs.coin_flip_envs", "Coin_InvB", "coin", "B", 2),
    ("coin_C1", "envs.coin_flip_envs", "Coin_InvC", "coin", "C", 1),
    ("coin_C3", "envs.coin_flip_envs", "Coin_InvC", "coin", "C", 3),
    ("dice_A1", "envs.dice_roll_envs", "Dice_InvA", "dice", "A", 1),
    ("dice_B2", "envs.dice_roll_envs", "Dice_InvB", "dice", "B", 2),
    ("dice_C1", "envs.dice_roll_envs", "Dice_InvC", "dice", "C", 1),
    ("gbm_A1", "envs.gbm_envs", "GBM_InvA", "gbm", "A", 1),
    ("gbm_A2", "envs.gbm_envs", "GBM_InvA", "gbm", "A", 2),
    ("gbm_B2", "envs.gbm_envs", "GBM_InvB", "gbm", "B", 2),
    ("gbm_C1", "envs.gbm_envs", "GBM_InvC", "gbm", "C", 1),
    ("dicesh_A", "envs.dice_roll_sh_envs", "Dice_SH_InvA", "dice_sh", "A", 1),
    ("dicesh_B", "envs.dice_roll_sh_envs", "Dice_SH_InvB", "dice_sh", "B", 1),
    ("dicesh_C", "envs.dice_roll_sh_envs", "Dice_SH_InvC", "dice_sh", "C", 1),
]
ENV_STEPS = 600


def env_inputs(name: str, action_dim: int, n_draws: int, family: str):
    """Seeded actions [T,A] and raw draws [T,n]: uniform(0,1) for the discrete families,
    standard normals for GBM (turned into returns by `env_returns`)."""
    seed = 1000 + sum(ord(ch) for ch in name)
    rs = np.random.RandomState(seed)
    actions = rs.uniform(-0.99, 0.99, size=(ENV_STEPS, action_dim))
    # saturated / degenerate actions exercise the exact-equality done flags
    for t in range(7, ENV_STEPS, 41):
        actions[t, -1] = 0.99
    for t in range(19, ENV_STEPS, 53):
        actions[t, :] = -0.99
    for t in range(29, ENV_STEPS, 67):
        actions[t, :] = 1e-7
    for t in range(3, ENV_STEPS, 5):
        actions[t] = np.abs(actions[t]) * 0.3    # calm stretches so episodes get long
    draws = rs.standard_normal((ENV_STEPS, n_draws)) if family == "gbm" else rs.random_sample((ENV_STEPS, n_draws))
    return actions, draws


def env_returns(family: str, draws: np.ndarray) -> np.ndarray:
    """Raw draws -> the return values the env would have sampled."""
    if family == "coin":
        return np.where(draws < 0.5, 0.5, -0.4)
    if family in ("dice", "dice_sh"):
        return np.where(draws < 1 / 6, 0.5, np.where(draws < 2 / 6, -0.5, 0.05))
    drift, vol = 0.0540025395205692, 0.1897916175617430
    return (drift - vol ** 2 / 2) + vol * draws
