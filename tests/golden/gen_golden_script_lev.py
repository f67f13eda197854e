"""
Runs the UNMODIFIED reference harness /root/reference/tests/test_script_lev.py (its
`__main__` body: all 11 `lev.lev_exp` functions at N = 1e4, H = 3e2, seed 420, CPU)
and records what it saves - tests/golden/script_lev.npz - so that the GPU test
(tests/test_script_lev_gpu.py) can run the same sequence of calls against the
injected engine module and compare.

matplotlib is absent here: `plotting.plots_multiverse` is replaced by a stub that
records its calls (the figures are outside the hot path); `np.save` is wrapped to
capture the arrays.  Nothing of the reference is copied: the file is executed
where it lies.

    python tests/golden/gen_golden_script_lev.py      (build container only)
"""
import hashlib
import os
import runpy
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

KEEP_EVERY = 13      # time columns kept of the per-step arrays (plus the last one)


def main():
    ref_shim.load("lev.lev_exp")            # installs the numpy / gym stubs, puts the tree on sys.path
    plot_calls = []
    plots = types.ModuleType("plotting.plots_multiverse")
    for name in ("plot_inv1", "plot_inv2", "plot_inv3", "plot_inv4"):
        setattr(plots, name, (lambda n: (lambda *a, **k: plot_calls.append((n, a[-1]))))(name))
    pkg = types.ModuleType("plotting")
    pkg.plots_multiverse = plots
    sys.modules["plotting"], sys.modules["plotting.plots_multiverse"] = pkg, plots

    saved = {}
    real_save = np.save

    def capture(path, arr, *a, **k):
        saved[os.path.basename(path)[:-4]] = np.array(arr)
        return real_save(path, arr, *a, **k)

    np.save = capture
    tmp = tempfile.mkdtemp()
    old = os.getcwd()
    try:
        # the harness writes ./results/... relative to the cwd and imports `tests.test_input_lev`, `lev.lev_exp`
        os.chdir(tmp)
        for name in ("tests", "lev"):
            sys.modules.pop(name, None)
        with ref_shim.quiet() as buf:
            runpy.run_path(os.path.join(ref_shim.REFERENCE_ROOT, "tests", "test_script_lev.py"), run_name="__main__")
    finally:
        os.chdir(old)
        np.save = real_save
    text = buf.getvalue()
    out = {"text": np.array(text), "plot_calls": np.array([f"{n}:{os.path.basename(p)}" for n, p in plot_calls])}
    for name, a in saved.items():
        if name.endswith("_val_T"):
            out[name + "_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest())
            out[name + "_sub"] = a[:, ::10]
        elif name.endswith("_lev"):
            out[name] = a
        else:
            cols = list(range(0, a.shape[-1], KEEP_EVERY))
            if cols[-1] != a.shape[-1] - 1:
                cols.append(a.shape[-1] - 1)
            out[name + "_cols"] = np.array(cols)
            out[name] = a[..., cols]
        print(name, a.shape, a.dtype)
    np.savez_compressed(os.path.join(HERE, "script_lev.npz"), **out)
    print(len(text.splitlines()), "printed lines;", [c for c in out["plot_calls"]])


if __name__ == "__main__":
    main()
