"""
Fixture generators that drive the UNMODIFIED reference env / replay classes with
injected random draws (see gen_golden.py for how to run).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
import golden_io  # noqa: E402


class _Feeder:
    """Stands in for np.random.choice / np.random.normal: hands out pre-drawn returns."""

    def __init__(self):
        self.queue = []

    def choice(self, a, p=None, size=None):
        v = self.queue.pop(0)
        assert len(v) == int(size)
        assert all(x in list(a) for x in v)
        return np.array(v, dtype=np.float64)

    def normal(self, loc=0.0, scale=1.0, size=None):
        v = self.queue.pop(0)
        assert len(v) == int(size)
        return np.array(v, dtype=np.float64)


class _Numpy122:
    """`np` as the reference module sees it, with numpy 1.22's reading of a ragged list handed to
    np.array: an element that is a 1-element array counts as its scalar (numpy >= 1.24 raises
    "inhomogeneous shape").  Generator-side only; the reference source is untouched."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def array(obj, *a, **k):
        if isinstance(obj, list) and any(isinstance(e, np.ndarray) and e.size == 1 and e.ndim > 0 for e in obj):
            obj = [e.reshape(()).item() if isinstance(e, np.ndarray) and e.size == 1 else e for e in obj]
        return np.array(obj, *a, **k)


def gen_env(only=None):
    for name, module, cls, family, investor, n_g in golden_io.ENV_CASES:
        if only is not None and name not in only:
            continue
        mod = ref_shim.load(module)
        if cls == "Dice_SH_INSURED":
            mod.np = _Numpy122()
        with ref_shim.reference_cwd():
            env = getattr(mod, cls)() if family == "dice_sh" else getattr(mod, cls)(n_g)
        a_dim = env.action_space.shape[0]
        n_draw = 1 if family == "dice_sh" else n_g
        actions, draws = golden_io.env_inputs(name, a_dim, n_draw, family)
        rets = golden_io.env_returns(family, draws)
        feeder = _Feeder()
        old_choice, old_normal = np.random.choice, np.random.normal
        np.random.choice, np.random.normal = feeder.choice, feeder.normal
        states, rewards, dones, risks, resets = [], [], [], [], []
        try:
            state0 = np.array(env.reset(), dtype=np.float64)
            for t in range(golden_io.ENV_STEPS):
                feeder.queue.append(list(rets[t]))
                ns, rew, done, risk = env.step(actions[t].copy())
                states.append(np.array(ns, dtype=np.float64).copy())
                rewards.append(float(rew))
                dones.append([bool(done[0]), bool(done[1])])
                risks.append(np.array(risk, dtype=np.float64).copy())
                resets.append(bool(done[0]))
                if done[0]:
                    env.reset()
        finally:
            np.random.choice, np.random.normal = old_choice, old_normal
            if cls == "Dice_SH_INSURED":
                mod.np = np
        out = os.path.join(HERE, f"env_{name}.npz")
        np.savez_compressed(out, actions=actions, returns=rets, state0=state0, states=np.array(states),
                            rewards=np.array(rewards), dones=np.array(dones), risks=np.array(risks),
                            resets=np.array(resets))
        print("wrote", out, np.array(states).shape, "episodes:", int(np.sum(resets)))


def gen_replay():
    """
    tools/replay_torch.py ReplayBufferTorch (and the NumPy twin tools/replay.py) driven
    through their public store_exp / sample_exp with `randperm` / `np.random.choice`
    answering the fixture's pre-drawn batches.  terminal_memory is zeroed after
    construction (T.empty leaves it uninitialised).
    """
    import torch

    mod = ref_shim.load("tools.replay_torch")
    for case in golden_io.REPLAY_CASES:
        st = golden_io.replay_stream(case)
        inputs = golden_io.replay_inputs_dict(case)
        inputs["gpu"] = "cpu"
        buf = mod.ReplayBufferTorch(inputs)
        buf.terminal_memory[:] = False
        pending = {"batch": None}
        real_randperm = torch.randperm

        def fake_randperm(n, *a, **k):
            b = pending["batch"]
            assert int(b.max()) < n
            return torch.as_tensor(b)

        out = {}
        ev = 0
        torch.randperm = fake_randperm
        try:
            for i in range(case["fill"] + 1):
                while ev < len(case["events"]) and case["events"][ev] == i:
                    pending["batch"] = st["batches"][ev]
                    s, a, r, s2, d, eff = buf.sample_exp()
                    out[f"ev{ev}_batch"] = st["batches"][ev]
                    out[f"ev{ev}_states"] = np.array(torch.as_tensor(s).numpy(), dtype=np.float32).copy()
                    out[f"ev{ev}_actions"] = np.array(torch.as_tensor(a).numpy(), dtype=np.float32).copy()
                    out[f"ev{ev}_rewards"] = np.array(torch.as_tensor(r).numpy(), dtype=np.float32).copy()
                    out[f"ev{ev}_next_states"] = s2.numpy().copy()
                    out[f"ev{ev}_dones"] = d.numpy().copy()
                    out[f"ev{ev}_eff"] = np.atleast_1d(torch.as_tensor(eff).numpy()).astype(np.int64)
                    ev += 1
                if i == case["fill"]:
                    break
                buf.store_exp(st["state"][i], st["action"][i], float(st["reward"][i]), st["next_state"][i],
                              bool(st["done"][i]))
        finally:
            torch.randperm = real_randperm
        assert ev == len(case["events"])
        out["mem_idx"] = np.array(buf.mem_idx)
        out["reward_memory"] = buf.reward_memory[: case["fill"]].numpy().copy()
        path = os.path.join(HERE, f"replay_{case['name']}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, "events:", ev, "episodes:", int(st["done"].sum()))


def gen_bigbrain():
    """coin_big_brain_lev / dice_big_brain_lev of the unmodified reference on seeded outcomes."""
    import torch

    lev = ref_shim.load("lev.lev_exp")
    dev = torch.device("cpu")
    for case in golden_io.BIGBRAIN_CASES:
        oc = golden_io.draw_outcomes(case)
        n, h = case["n"], case["h"]
        up, dn = case["up_r"], case["down_r"]
        # lev/coin_flip.py:139-152 verbatim in meaning
        value_0 = torch.tensor(case["v0"], device=dev)
        investors = torch.tensor(int(n), dtype=torch.int32, device=dev)
        horizon = torch.tensor(int(h), dtype=torch.int32, device=dev)
        asym = torch.tensor(1e-12, device=dev)
        bigger = np.abs(dn) if np.abs(up) >= np.abs(dn) else -np.abs(up)
        lev_factor = torch.tensor(1 / bigger, device=dev)
        lev_factor = lev_factor - asym if np.abs(up) > np.abs(dn) else lev_factor + asym
        assert float(lev_factor) == golden_io.bigbrain_lev_factor(case) and lev_factor.dtype == torch.float64
        with ref_shim.quiet() as buf:
            if case["kind"] == "coin":
                data = lev.coin_big_brain_lev(dev, torch.tensor(oc.astype(np.float32)), investors, horizon,
                                              case["top"], value_0, up, dn, lev_factor, *case["stop"], *case["roll"])
            else:
                data = lev.dice_big_brain_lev(dev, torch.tensor(oc.astype(np.int64)), investors, horizon,
                                              case["top"], value_0, up, dn, case["mid_r"], lev_factor,
                                              *case["stop"], *case["roll"])
        data = data.numpy()
        cols = golden_io.kept_columns(case)
        out = os.path.join(HERE, f"bigbrain_{case['name']}.npz")
        np.savez_compressed(out, data=data[:, :, :, cols], cols=np.asarray(cols, dtype=np.int64),
                            text=np.array(buf.getvalue()))
        print("wrote", out, data.shape, os.path.getsize(out) // 1024, "KiB")

    # Kelly table (coin_galaxy_brain_lev, lev/lev_exp.py:455-505) on a reduced grid
    with ref_shim.quiet():
        table = lev.coin_galaxy_brain_lev(dev, *golden_io.GALAXY_GRID)
    out = os.path.join(HERE, "galaxy_brain.npz")
    np.savez_compressed(out, data=table.numpy())
    print("wrote", out, tuple(table.shape))
