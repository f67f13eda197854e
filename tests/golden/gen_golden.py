"""
Generates the committed golden fixtures in tests/golden/ by running the
UNMODIFIED reference (/root/reference) on seeded, pre-drawn inputs.

Run in the build container only (the reference tree does not travel):

    python tests/golden/gen_golden.py [lev|env|replay|all]

Inputs are never stored: each fixture records the seed and the SHA-256 of the
regenerated input bytes (numpy RandomState streams are frozen, NEP 19);
`tests/golden_io.py` regenerates the inputs and checks the hash.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
import golden_io  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ------------------------------------------------------------------------- lev
def gen_lev():
    import torch as T

    lx = ref_shim.load("lev.lev_exp")
    dev = T.device("cpu")
    for case in golden_io.LEV_CASES:
        name, kind = case["name"], case["kind"]
        n, h, top, v0 = case["n"], case["h"], case["top"], case["v0"]
        outcomes = golden_io.draw_outcomes(case)
        args_common = (T.tensor(n, dtype=T.int32), T.tensor(h, dtype=T.int32), top, T.tensor(v0))
        grid = case["grid"]
        with ref_shim.quiet():
            if kind == "coin":
                oc = T.tensor(outcomes.astype(np.float32))
                data, data_T = lx.coin_smart_lev(dev, oc, *args_common, case["up_r"], case["down_r"], *grid)
            elif kind == "dice":
                oc = T.tensor(outcomes.astype(np.int64))
                data, data_T = lx.dice_smart_lev(dev, oc, *args_common, case["up_r"], case["down_r"],
                                                 case["mid_r"], *grid)
            elif kind == "dice_sh":
                oc = T.tensor(outcomes.astype(np.int64))
                data, data_T = lx.dice_sh_smart_lev(dev, oc, *args_common, case["up_r"], case["down_r"],
                                                    case["mid_r"], *case["sh"], *grid)
            elif kind == "gbm":
                oc = T.tensor(outcomes)
                data, data_T = lx.gbm_smart_lev(dev, oc, *args_common, *grid)
            else:
                raise ValueError(kind)
        out = os.path.join(HERE, f"lev_{name}.npz")
        # per-step statistics are kept in full only for small cases; the big
        # ones keep a strided subset of time columns to stay small in git
        data = data.numpy()
        cols = golden_io.kept_columns(case)
        np.savez_compressed(
            out,
            input_sha256=np.array(sha(outcomes)),
            data=data[:, :, cols],
            cols=np.asarray(cols, dtype=np.int64),
            data_T=data_T.numpy(),
        )
        print("wrote", out, data.shape, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("lev", "all"):
        gen_lev()
    if what in ("bigbrain", "all") and hasattr(golden_io, "BIGBRAIN_CASES"):
        from gen_golden_more import gen_bigbrain
        gen_bigbrain()
    if what in ("env", "all") and hasattr(golden_io, "ENV_CASES"):
        from gen_golden_more import gen_env
        gen_env()
    if what in ("replay", "all") and hasattr(golden_io, "REPLAY_CASES"):
        from gen_golden_more import gen_replay
        gen_replay()
