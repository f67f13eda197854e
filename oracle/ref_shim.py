"""
TEST INFRASTRUCTURE ONLY - loader for the *unmodified* reference tree.

The reference (majidsina/rlmd) is pure Python. It cannot travel to the GPU box,
so it is imported only in the build container (where /root/reference exists) by
`tests/golden/gen_golden.py` to produce the committed golden fixtures and by
`tests/test_oracle_vs_reference.py` (skipped when the tree is absent).

Nothing in rlmd_b200/ may import this module.

Why a shim is needed: the reference pins
numpy 1.22 / gym 0.24; under numpy 2.x `np.float_` and `np.bool8` are gone and
`gym` is not installed.  We alias the two numpy names and register a stub `gym`
that exposes only what envs/*.py touch (`gym.Env`, `gym.spaces.Box`).  No
reference source is modified or copied.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("RLMD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lev", "lev_exp.py"))


def _install_stubs() -> None:
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if not hasattr(np, "bool8"):
        np.bool8 = np.bool_
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        spaces = types.ModuleType("gym.spaces")

        class Env:  # noqa: D401 - minimal stand-in
            pass

        class Box:
            def __init__(self, low, high, shape, dtype):
                self.low = np.full(shape, low, dtype)
                self.high = np.full(shape, high, dtype)
                self.shape = shape
                self.dtype = dtype

            def sample(self):
                return np.random.uniform(self.low, self.high)

        gym.Env, spaces.Box, gym.spaces = Env, Box, spaces
        sys.modules["gym"], sys.modules["gym.spaces"] = gym, spaces


@contextlib.contextmanager
def reference_cwd():
    """The reference modules do `sys.path.append("./")`: run from its root."""
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def load(module: str):
    """Import `module` (e.g. "lev.lev_exp") from the reference tree."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _install_stubs()
    sys.dont_write_bytecode = True  # the tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    with reference_cwd():
        # our own `tests`/`tools` packages must not shadow the reference's
        for name in ("tests", "tools", "envs", "lev"):
            mod = sys.modules.get(name)
            if mod is not None and not str(getattr(mod, "__file__", "")).startswith(
                REFERENCE_ROOT
            ):
                del sys.modules[name]
        return importlib.import_module(module)


@contextlib.contextmanager
def quiet():
    """The lev_exp functions print three lines per leverage."""
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf
