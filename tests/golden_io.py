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
    ("coin_A8", "envs.coin_flip_envs", "Coin_InvA", "coin", "A", 8),     # n >= 8: NumPy's pairwise np.sum order
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
    # Dice_SH_INSURED.step builds `risk` from scalars mixed with 1-element arrays (envs/dice_roll_sh_envs.py:
    # 178-179, 228-231): numpy 1.22 (the reference's pin) takes such elements as scalars, numpy 2 refuses.  The
    # generator hands the reference module a numpy proxy whose `array` does what 1.22 did (gen_golden_more.py).
    ("dicesh_I", "envs.dice_roll_sh_envs", "Dice_SH_INSURED", "dice_sh", "I", 1),
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


# -------------------------------------------------------------------- replay
# Seeded transition streams + sampling events for tools/replay_torch.py.
# `lens`: episode lengths in order (cycled); `fill`: transitions stored in total;
# `events`: fill levels at which a mini-batch is sampled (before storing further).
REPLAY_CASES = [
    # plain gather (multi_steps == 1)
    dict(name="n1", mem=600, s=5, a=1, n=1, gamma=0.99, dyna="M", batch=64, r0=None,
         lens=[7, 12, 5, 30, 9], fill=500, events=[64, 130, 500], seed=11),
    # additive n-step, sampling while the first episode is still running and afterwards
    dict(name="n3_A", mem=400, s=5, a=1, n=3, gamma=0.5, dyna="A", batch=32, r0=None,
         lens=[40, 3, 5, 2, 17, 1, 1, 9], fill=330, events=[33, 40, 41, 77, 200, 330], seed=12),
    # multiplicative n-step (the multiplicative envs' setting), reward floor active
    dict(name="n5_M", mem=700, s=5, a=1, n=5, gamma=0.99, dyna="M", batch=48, r0=0.995,
         lens=[9, 33, 4, 6, 60, 5, 2, 11], fill=640, events=[48, 49, 100, 333, 640], seed=13),
    # n = 10, wider state / action (Dice_SH_InvC: 6 / 3), first transition terminal
    dict(name="n10_M", mem=900, s=6, a=3, n=10, gamma=0.97, dyna="M", batch=96, r0=None,
         lens=[1, 14, 25, 1, 8, 58, 10, 3], fill=800, events=[96, 97, 250, 611, 800], seed=14),
    dict(name="n10_A", mem=500, s=3, a=2, n=10, gamma=0.9, dyna="A", batch=40, r0=-0.5,
         lens=[12, 12, 30, 2, 21], fill=450, events=[40, 45, 222, 450], seed=15),
    # n = 2: the shortest multi-step history
    dict(name="n2_M", mem=300, s=5, a=1, n=2, gamma=0.99, dyna="M", batch=32, r0=None,
         lens=[5, 6, 7, 1, 2, 50], fill=280, events=[32, 150, 280], seed=16),
]


def replay_case(name: str) -> dict:
    for c in REPLAY_CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)


def replay_stream(case: dict) -> dict:
    """The transitions of a case: float64 arrays as an env would hand them over, python-bool dones."""
    rs = np.random.RandomState(case["seed"])
    f, s, a = case["fill"], case["s"], case["a"]
    done = np.zeros(f, dtype=bool)
    pos, k = -1, 0
    while True:
        pos += case["lens"][k % len(case["lens"])]
        k += 1
        if pos >= f:
            break
        done[pos] = True
    state = rs.standard_normal((f, s))
    action = rs.uniform(-0.99, 0.99, size=(f, a))
    next_state = rs.standard_normal((f, s))
    if case["dyna"] == "A":
        reward = rs.standard_normal(f)
    else:
        reward = 1.0 + 0.01 * rs.standard_normal(f)
    batches = [rs.permutation(e)[: case["batch"]].astype(np.int64) for e in case["events"]]
    return dict(state=state, action=action, reward=reward, next_state=next_state, done=done, batches=batches)


def replay_inputs_dict(case: dict) -> dict:
    """The `inputs` dictionary keys tools/replay_torch.py:64-82 reads."""
    return {
        "gpu": "cuda:0", "input_dims": (case["s"],), "num_actions": case["a"], "mini_batch_size": case["batch"],
        "discount": case["gamma"], "multi_steps": case["n"], "r_abs_zero": case["r0"], "dynamics": case["dyna"],
        "buffer": case["mem"], "n_cumsteps": case["mem"] + 100,
    }


# ------------------------------------------------------------------ big brain
# coin_big_brain_lev / dice_big_brain_lev (lev/lev_exp.py:270-452, 741-932): stop-loss
# x retention grids (low, high, incr) as the scripts pass them; lev_factor is built
# the way lev/coin_flip.py:146-152 / lev/dice_roll.py:133-139 build it.
BIGBRAIN_CASES = [
    # the reference's own smoke-test scale and grids (tests/test_script_lev.py:61-66,102-106)
    dict(name="coin_inv2_testscale", kind="coin", n=10000, h=300, top=1, v0=1e2, seed=430, up_r=0.5, down_r=-0.4,
         p=(0.5,), stop=(0.10, 0.10, 0.10), roll=(0.00, 0.00, 0.10), stride=23),
    dict(name="coin_inv3_testscale", kind="coin", n=10000, h=300, top=1, v0=1e2, seed=431, up_r=0.5, down_r=-0.4,
         p=(0.5,), stop=(0.70, 0.80, 0.10), roll=(0.70, 0.80, 0.10), stride=23),
    dict(name="dice_inv2_testscale", kind="dice", n=10000, h=300, top=1, v0=1e2, seed=432, up_r=0.5, down_r=-0.5,
         mid_r=0.05, p=(1 / 6, 1 / 6), stop=(0.10, 0.10, 0.10), roll=(0.00, 0.00, 0.10), stride=23),
    dict(name="dice_inv3_testscale", kind="dice", n=10000, h=300, top=1, v0=1e2, seed=433, up_r=0.5, down_r=-0.5,
         mid_r=0.05, p=(1 / 6, 1 / 6), stop=(0.70, 0.80, 0.10), roll=(0.70, 0.80, 0.10), stride=23),
    # top-K > 1, ragged sizes, mixed grids (retention 0 and > 0 in one call)
    dict(name="coin_grid_top4", kind="coin", n=1237, h=45, top=4, v0=1e2, seed=434, up_r=0.5, down_r=-0.4,
         p=(0.5,), stop=(0.05, 0.95, 0.30), roll=(0.00, 0.90, 0.45), stride=1),
    dict(name="dice_grid_top3", kind="dice", n=1111, h=37, top=3, v0=1e2, seed=435, up_r=0.5, down_r=-0.5,
         mid_r=0.05, p=(1 / 6, 1 / 6), stop=(0.10, 0.90, 0.40), roll=(0.00, 0.90, 0.45), stride=1),
    # the down move is the bigger payoff: lev_factor < 0 branch of the scripts
    dict(name="coin_down_heavy", kind="coin", n=640, h=30, top=2, v0=1e2, seed=436, up_r=0.3, down_r=-0.5,
         p=(0.6,), stop=(0.20, 0.60, 0.20), roll=(0.50, 0.50, 0.10), stride=1),
]


def bigbrain_case(name: str) -> dict:
    for c in BIGBRAIN_CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)


def bigbrain_lev_factor(case: dict):
    """float64 value of the scripts' LEV_FACTOR (a float64 0-dim tensor there)."""
    up, dn = case["up_r"], case["down_r"]
    bigger = np.abs(dn) if np.abs(up) >= np.abs(dn) else -np.abs(up)
    asym = float(np.float32(1e-12))          # T.tensor(1e-12) is fp32
    return float(1 / bigger) - asym if np.abs(up) > np.abs(dn) else float(1 / bigger) + asym


GALAXY_GRID = (0.2, 0.8, 0.05, 0.2, 0.8, 0.05, 0.25, 0.75, 0.25)   # ru, rd, pu (low, high, incr)


# -------------------------------------------------------------------- market
# (fixture name, investor, history (Dx), n_assets, obs_days, time_length, steps)
MARKET_CASES = [
    ("A_D1_n1", "A", False, 1, 1, 40, 400),
    ("B_D1_n3", "B", False, 3, 1, 60, 400),
    ("C_D1_n9", "C", False, 9, 1, 50, 400),        # n >= 8: NumPy's pairwise summation order
    ("A_Dx_n3_d5", "A", True, 3, 5, 45, 300),
    ("B_Dx_n1_d3", "B", True, 1, 3, 30, 300),
    ("C_Dx_n26_d4", "C", True, 26, 4, 40, 120),
]


def market_case(name: str):
    for c in MARKET_CASES:
        if c[0] == name:
            return c
    raise KeyError(name)


def market_inputs(case):
    """
    Seeded inputs of one market fixture: a price history [L, n] (geometric random
    walk, 2 % daily volatility) and actions [T, A]; a few actions are saturated /
    vanishing so that every termination rule fires.
    """
    name, investor, history, n, d, tl, steps = case
    seed = int(hashlib.sha256(("market_" + name).encode()).hexdigest()[:8], 16)
    rs = np.random.RandomState(seed)
    a_dim = {"A": 0, "B": 1, "C": 2}[investor] + n
    length = 4000
    prices = 100.0 * np.exp(np.cumsum(0.02 * rs.standard_normal((length, n)), axis=0))
    actions = rs.uniform(-0.99, 0.99, size=(steps, a_dim))
    special = rs.random_sample(steps)
    actions[special < 0.02] = 0.99
    actions[(special >= 0.02) & (special < 0.04)] = 1e-8
    return prices, actions


def market_observed(extract: np.ndarray, time_step: int, obs_days: int) -> np.ndarray:
    """tools/env_resources.py:203-226 with action_days = 1."""
    if obs_days == 1:
        return extract[time_step]
    hi = time_step + obs_days if time_step > 0 else obs_days
    return extract[time_step:hi].reshape(-1)[::-1]


def market_episode_start(k: int, length: int, time_length: int, obs_days: int) -> int:
    """First row of the k-th episode's market extract (deterministic stand-in for time_slice)."""
    return (977 * k) % (length - time_length - 2 * obs_days - 8)
