"""TEST INFRASTRUCTURE ONLY -- import shim for the reference's pure-Python engine.

Lets `import uxsim` work from /root/reference in THIS container (no matplotlib, missing font blob),
without touching the reference tree.  Used only by the golden-vector generator
(`tests/golden/make_golden.py`) and by container-only cross-checks; nothing on the product path,
and nothing that runs on the GPU box, may import this (the reference does not travel).

What it patches:
  * registers stub modules for matplotlib{,.pyplot,.lines,.font_manager} in sys.modules
    (font_manager.findSystemFonts -> [] so utils.get_font_for_matplotlib terminates);
  * monkey-patches uxsim.analyzer.load_font_data to read a font that exists in the tree.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("UXSIM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "uxsim"))


def import_reference():
    """Return the reference `uxsim` package (pure-Python engine)."""
    if not reference_available():
        raise ImportError(f"reference not present at {REFERENCE_ROOT}")
    if "uxsim" in sys.modules and getattr(sys.modules["uxsim"], "_b200_shimmed", False):
        return sys.modules["uxsim"]

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, name):
            return _Anything()

        def __iter__(self):
            return iter(())

        def __setitem__(self, k, v):
            pass

        def __getitem__(self, k):
            return _Anything()

    def _stub(name):
        m = types.ModuleType(name)
        def _attr(attr):  # dunder lookups (__file__, __path__, __spec__ ...) must fail like on a real attribute-less module:
            if attr.startswith("__"):  # other packages (torch) walk sys.modules and inspect them
                raise AttributeError(attr)
            return _Anything()

        m.__getattr__ = _attr  # type: ignore[attr-defined]
        return m

    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.pyplot.rcParams = {}
        mpl.lines = _stub("matplotlib.lines")
        mpl.font_manager = _stub("matplotlib.font_manager")
        mpl.font_manager.findSystemFonts = lambda *a, **k: []
        mpl.colors = _stub("matplotlib.colors")
        mpl.cm = _stub("matplotlib.cm")
        for k, v in {
            "matplotlib": mpl,
            "matplotlib.pyplot": mpl.pyplot,
            "matplotlib.lines": mpl.lines,
            "matplotlib.font_manager": mpl.font_manager,
            "matplotlib.colors": mpl.colors,
            "matplotlib.cm": mpl.cm,
        }.items():
            sys.modules[k] = v

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import uxsim  # noqa: E402
    import uxsim.analyzer as analyzer  # noqa: E402

    def _load_font_data(font_path=None):
        for cand in ("Inconsolata.otf",):
            p = os.path.join(REFERENCE_ROOT, "uxsim", "files", cand)
            if os.path.exists(p):
                with open(p, "rb") as f:
                    return f.read()
        return b""

    analyzer.load_font_data = _load_font_data
    uxsim._b200_shimmed = True
    return uxsim
