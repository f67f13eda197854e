"""
CPU-side checks of the drop-in boundary: the shared library loads without a
GPU and exports exactly the entry points include/rlmd_b200.h declares, the
ctypes mirror of the descriptor structs has the C layout, and compute calls
fail loudly (no CPU fallback) when no CUDA device is present.
"""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rlmd_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rlmd_b200 import _lib

    names = declared_functions()
    assert len(names) >= 8
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes signature in rlmd_b200/_lib.py"
    for n in _lib.EXPORTS:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_version_and_error_string():
    from rlmd_b200 import _lib

    assert _lib.lib.b200_version() >= 100
    assert isinstance(_lib.lib.b200_last_error(), bytes)


def test_struct_layout_matches_c(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof."""
    import subprocess

    from rlmd_b200 import _lib

    structs = {"b200_lev_desc": _lib.LevDesc}
    names = {"EnvDesc": "b200_env_desc", "ReplayDesc": "b200_replay_desc", "PeerSet": "b200_peer_set",
             "MarketDesc": "b200_market_desc", "CollectDesc": "b200_collect_desc", "BigBrainDesc": "b200_bigbrain_desc",
             "TallyPlan": "b200_tally_plan", "TallyPeers": "b200_tally_peers"}
    for extra, cname in names.items():
        if hasattr(_lib, extra):
            structs[cname] = getattr(_lib, extra)
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct)
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np

    from rlmd_b200 import engine
    from rlmd_b200._lib import B200Error

    with pytest.raises(B200Error):
        engine.lev_sweep("discrete", np.ones((2, 2), np.float32), 100.0, n_investors=4, horizon=4,
                         probs=(0.5, 0.5))


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (or the reference)."""
    pkg = os.path.join(ROOT, "rlmd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "ref_shim" not in text and "/root/reference" not in text, f


def test_param_range_matches_oracle_table():
    from oracle import lev_oracle as lo
    from rlmd_b200 import lev_exp

    for args in [(0.05, 1.0, 0.05), (0.1, 1.0, 0.1), (0.05, 0.95, 0.05), (0.7, 0.95, 0.05), (0.45, 0.95, 0.05),
                 (0.1, 0.1, 0.1), (0.0, 0.0, 0.1), (0.73, 1.0, 0.03), (-1.0, 1.0, 0.2), (0.4, 4.0, 0.4),
                 (0.2, 2.0, 0.2), (0.5, 1.0, 0.1), (0.7, 0.8, 0.1), (0.2, 0.8, 0.001)]:
        assert lev_exp.param_range(*args) == lo.param_range(*args), args


def test_factor_tables_match_oracle():
    import numpy as np

    from oracle import lev_oracle as lo
    from rlmd_b200 import lev_exp

    lev = lo.lev_grid(0.05, 1.0, 0.05)
    assert np.array_equal(lev_exp.coin_factor_table(lev, 0.5, -0.4), lo.coin_factors(lev, 0.5, -0.4))
    assert np.array_equal(lev_exp.dice_factor_table(lev, 0.5, -0.5, 0.05), lo.dice_factors(lev, 0.5, -0.5, 0.05))
    lev = lo.lev_grid(0.73, 1.0, 0.03)
    assert np.array_equal(lev_exp.dice_sh_factor_table(lev, 0.5, -0.5, 0.05, -1, 5, -1),
                          lo.dice_sh_factors(lev, 0.5, -0.5, 0.05, -1, 5, -1))


def test_tally_plan_sizes():
    """Workspace arithmetic of the tally path (host code only): sizes grow with the plan, the exchange
    buffer exists only across GPUs, bad plans are refused with a message."""
    import ctypes as C

    from rlmd_b200 import _lib

    def plan(rows, bins, grid=64, world=1):
        p = _lib.TallyPlan()
        p.rows_cap, p.bins_cap, p.grid_cap, p.world = rows, bins, grid, world
        return p

    small = _lib.lib.b200_tally_workspace_bytes(C.byref(plan(1000, 1000)))
    big = _lib.lib.b200_tally_workspace_bytes(C.byref(plan(1_000_000, 1_000_000)))
    multi = _lib.lib.b200_tally_workspace_bytes(C.byref(plan(1_000_000, 1_000_000, world=8)))
    assert 0 < small < big < multi
    # table (12 B x 2^21) + bins (12 B) + wealth buffer (64 x 4 B) per possible tuple
    assert big >= (1 << 21) * 12 + 1_000_000 * (12 + 256)
    assert _lib.lib.b200_tally_exchange_bytes(C.byref(plan(1000, 1000))) == 0
    assert _lib.lib.b200_tally_exchange_bytes(C.byref(plan(1_000_000, 1_000_000, world=8))) >= 2 * 16 * (1 << 20)
    assert _lib.lib.b200_tally_workspace_bytes(C.byref(plan(0, 10))) < 0
    assert b"rows_cap" in _lib.lib.b200_last_error()
    assert _lib.lib.b200_tally_workspace_bytes(C.byref(plan(10, 10, grid=65))) < 0
    assert _lib.lib.b200_tally_workspace_bytes(C.byref(plan(10, 10, world=9))) < 0


def test_constants_match_the_header(tmp_path):
    """The Python mirror of the header's constants (flags, limits, enums) is what a C compiler sees."""
    import subprocess

    from rlmd_b200 import _lib, sharding

    pairs = {
        "B200_MAX_PEERS": _lib.MAX_PEERS, "B200_MAX_GRID": _lib.MAX_GRID, "B200_MAX_OUTCOMES": _lib.MAX_OUTCOMES,
        "B200_LEV_DISCRETE": _lib.LEV_DISCRETE, "B200_LEV_GBM": _lib.LEV_GBM, "B200_SRC_STREAM": _lib.SRC_STREAM,
        "B200_SRC_PHILOX": _lib.SRC_PHILOX, "B200_MODE_CHAIN": _lib.MODE_CHAIN, "B200_MODE_LOG": _lib.MODE_LOG,
        "B200_LEV_FLAG_FINAL_ONLY": _lib.LEV_FLAG_FINAL_ONLY, "B200_LEV_FLAG_STATE_OUT": _lib.LEV_FLAG_STATE_OUT,
        "B200_LEV_FLAG_BESIDE_SWEEP": _lib.LEV_FLAG_BESIDE_SWEEP, "B200_TALLY_INFO_WORDS": _lib.TALLY_INFO_WORDS,
        "B200_DT_U8": _lib.DT_U8, "B200_DT_I32": _lib.DT_I32, "B200_DT_I64": _lib.DT_I64, "B200_DT_F32": _lib.DT_F32,
        "B200_DT_F64": _lib.DT_F64, "B200_ENV_MAX_GAMBLES": _lib.ENV_MAX_GAMBLES,
        "B200_REPLAY_MAX_STEPS": _lib.REPLAY_MAX_STEPS, "B200_MARKET_MAX_ASSETS": _lib.MARKET_MAX_ASSETS,
        "B200_BB_MAX_POINTS": _lib.BB_MAX_POINTS,
        "B200_PEER_FLAG_ERROR_WORD": sharding.PeerWorkspace.ERROR_WORD,
        "B200_PEER_FLAG_WORDS": 2 * sharding.PeerWorkspace.FLAG_WORDS,      # uint32 words in 8-byte slots
    }
    lines = ['#include <stdio.h>', f'#include "{HEADER}"', "int main(void){"]
    lines += [f'printf("{name} %lld\\n", (long long)({name}));' for name in pairs]
    lines.append("return 0;}")
    src = tmp_path / "consts.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "consts"
    subprocess.check_call(["gcc", "-std=c99", "-o", str(exe), str(src)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for name, want in pairs.items():
        assert int(out[name]) == int(want), (name, out[name], want)
