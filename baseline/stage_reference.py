"""Stage the UNMODIFIED reference Python package under baseline/_ref/ (git-ignored, but it travels to the GPU box with the
snapshot) so that bench.py can time the reference's pure-Python engine on the GPU box's own host cores.

    python baseline/stage_reference.py

`pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>`
was tried first and fails in this image: the package's build backend (scikit-build-core) configures the nanobind C++ extension
and nanobind is not installed (CMake: "Could not find a package configuration file provided by nanobind").  The pure-Python
engine needs no build, so this script does what the wheel would do for it: it copies the `uxsim` package tree (Python files and
the small data files they read) as it is.  Nothing is edited; nothing is copied into the tracked tree.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("UXSIM_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage() -> bool:
    pkg = os.path.join(SRC, "uxsim")
    if not os.path.isdir(pkg):
        return False
    out = os.path.join(DST, "uxsim")
    if os.path.isdir(out):
        shutil.rmtree(out)
    keep_ext = {".py", ".otf", ".ttf", ".txt", ".csv", ".json"}

    def ignore(d, names):
        drop = []
        for n in names:
            p = os.path.join(d, n)
            if os.path.isdir(p):
                if n in ("trafficpp", "__pycache__", "ResultGUIViewer"):
                    drop.append(n)
            elif os.path.splitext(n)[1].lower() not in keep_ext or os.path.getsize(p) > (4 << 20):
                drop.append(n)
        return drop

    shutil.copytree(pkg, out, ignore=ignore)
    # the reference's own verification tests and the three small CSVs they read (tests/test_reference_suite.py runs them, unmodified,
    # against this backend on the GPU box: the `--cuda` counterpart of the reference's `pytest --cpp`)
    tdir = os.path.join(DST, "tests")
    if os.path.isdir(tdir):
        shutil.rmtree(tdir)
    os.makedirs(tdir)
    for n in sorted(os.listdir(os.path.join(SRC, "tests"))):
        if n in ("conftest.py", "test_cpp_mode.py") or (n.startswith("test_verification_") and n.endswith(".py")):
            shutil.copy2(os.path.join(SRC, "tests", n), os.path.join(tdir, n))
    ddir = os.path.join(DST, "dat")
    os.makedirs(ddir, exist_ok=True)
    for n in ("siouxfalls_nodes.csv", "siouxfalls_links.csv", "siouxfalls_demand.csv"):
        if os.path.exists(os.path.join(SRC, "dat", n)):
            shutil.copy2(os.path.join(SRC, "dat", n), os.path.join(ddir, n))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{SRC} (uxsim package tree copied unmodified by baseline/stage_reference.py; pip install fails: nanobind absent)\n")
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else f"reference not found at {SRC}", DST)
    sys.exit(0 if ok else 1)
