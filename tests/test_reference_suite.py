"""The reference's OWN verification suite under this backend (SURVEY.md section 4 tier ii; VERDICT r1 "missing" #6).

The reference runs its test files against its C++ engine with `pytest --cpp` (/root/reference/tests/conftest.py:13-47).  Here the
same UNMODIFIED files -- straight road and bottlenecks, spillback, signals, merges / diverges, multilane, route choice, exceptional
cases, Sioux Falls statistics, the day-to-day DTA solvers -- run with every `World(...)` served by `uxsim_b200.World(cuda=True)` (tests/ref_suite_plugin.py),
and the files' own known-answer assertions decide.

  * CPU (`-m "not gpu"`): over the host emulation of the kernels, test files read from /root/reference/tests.
  * `-m gpu`: over libuxsim_b200.so on the B200, files from the staged copy baseline/_ref (baseline/stage_reference.py; git-ignored,
    it travels with the snapshot); skipped when nothing is staged.

EXPECTED_FAIL lists the tests that cannot pass in ANY array-backed mode and says why: they use Python-list semantics on results
that the reference's own C++ wrapper returns as numpy arrays too (uxsim_cpp_wrapper.py:292-333, 1309-1367), so `pytest --cpp`
fails them the same way.  Stochastic tests carry the reference's `flaky` rerun marker; a failing test is re-run up to three times
here as well (pytest-rerunfailures is not installed)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["test_verification_straight_road.py", "test_verification_node.py", "test_verification_multilane.py",
         "test_verification_route_choice.py", "test_verification_exceptional.py", "test_verification_sioux_falls.py",
         "test_verification_dta_solvers.py",  # (the plugin serves SolverDUE / SolverDSO_D2D by uxsim_b200.dta, as `cpp=True` selects the C++ adapter)
         "test_cpp_mode.py"]  # the reference's own C++-mode suite: every test of the files above once more with World(cpp=True), plus C++-mode specifics
# tests of a file that cannot apply here: they start NEW python processes that `import uxsim` and run example scripts / notebooks
DESELECT = {"test_cpp_mode.py": "not test_cpp_example_runs and not notebook"}
EXPECTED_FAIL = {
    # `["log_t"] + veh.log_t`: list concatenation with a per-vehicle log; CppVehicle.log_t is a numpy array as well
    "test_verification_straight_road.py": {"test_iterative_exec_rigorous", "test_iterative_exec_rigorous_random_size_old",
                                           "test_iterative_exec_rigorous_random_size_duration_t2"},
    # `assert cum1 == cum2` on two cumulative curves: ambiguous truth value of an array; CppLink.cum_arrival is a numpy array as well
    "test_verification_node.py": {"test_override_signal"},
    "test_cpp_mode.py": {
        # expects save_scenario / load_scenario / copy / save to raise NotImplementedError as in C++ mode; load_scenario IS implemented here
        # (SURVEY.md 8f-3), so the call goes on to open the file "dummy"
        "test_coverage_misc_world_methods",
        # expects addNode(existing name, auto_rename=True) to raise: the C++ wrapper renames on the Python side but registers the ORIGINAL name
        # with the C++ world, which rejects it (uxsim_cpp_wrapper.py:62-75); here auto_rename renames, as in Python mode
        "test_coverage_node_signal_and_rename",
    },
}


# Tests whose thresholds the reference's OWN engine misses in about half of its runs (they draw unseeded random numbers): not required
# to pass here either.  test_DTA_dynamic_congestion_pricing_on_highway_bottleneck asserts `average_travel_time_between(500, 3000) > 300`
# and `... (4000, 6000) < 200` after 40 day-to-day iterations; four runs of the reference's Python engine and solver in the dev container
# gave 285 / 292 / 301 / 341 and 199 / 175 / 190 / 221, six runs of this backend 273 ... 385 and 164 ... 174.
UNSTABLE_IN_REFERENCE = {"test_verification_dta_solvers.py": {"test_DTA_dynamic_congestion_pricing_on_highway_bottleneck"},
                         "test_cpp_mode.py": {"test_DTA_dynamic_congestion_pricing_on_highway_bottleneck"}}


def _reference_root():
    for r in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(r, "uxsim")) and os.path.isdir(os.path.join(r, "tests")):
            return r
    return None


def _run(ref_root, fname, lib, extra=()):
    env = dict(os.environ, UXSIM_B200_SUITE_LIB=lib, UXSIM_REFERENCE_ROOT=ref_root,
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT, os.environ.get("PYTHONPATH", "")]))
    if fname in DESELECT and not extra:
        extra = ("-k", DESELECT[fname])
    cmd = [sys.executable, "-m", "pytest", "-p", "ref_suite_plugin", "-p", "no:cacheprovider", "-q", "-rfEp", "--tb=line",
           os.path.join("tests", fname), *extra]
    out = subprocess.run(cmd, cwd=ref_root, env=env, capture_output=True, text=True, timeout=3000).stdout  # cwd: the tests read "dat/..."
    res = {}
    for m in re.finditer(r"^(PASSED|FAILED|ERROR) \S*?::(\w+)", out, re.M):
        res[m.group(2)] = m.group(1)
    return res, out


def _check(ref_root, fname, lib):
    res, out = _run(ref_root, fname, lib)
    assert res, f"no test results parsed from {fname}:\n{out[-2000:]}"
    expected = EXPECTED_FAIL.get(fname, set())
    unstable = UNSTABLE_IN_REFERENCE.get(fname, set())
    bad = sorted(t for t, r in res.items() if r != "PASSED" and t not in expected and t not in unstable)
    for _ in range(5):  # the reference marks its stochastic tests `flaky(reruns=...)`, up to 10
        if not bad:
            break
        again, out = _run(ref_root, fname, lib, ("-k", " or ".join(bad)))
        bad = sorted(t for t in bad if again.get(t) != "PASSED")
    assert not bad, f"{fname}: {bad} fail under the CUDA backend\n{out[-3000:]}"
    unexpected_pass = sorted(t for t in expected if res.get(t) == "PASSED")
    assert not unexpected_pass, f"{fname}: {unexpected_pass} are listed as expected failures but pass -- update EXPECTED_FAIL"
    return sum(1 for r in res.values() if r == "PASSED"), len(res)


@pytest.mark.parametrize("fname", FILES)
def test_reference_suite_emulated(fname, emu_api):
    ref_root = _reference_root()
    if ref_root is None:
        pytest.skip("the reference's tests are neither at /root/reference nor staged under baseline/_ref")
    pytest.importorskip("pandas")
    _check(ref_root, fname, "emu")


@pytest.mark.gpu
@pytest.mark.parametrize("fname", FILES)
def test_reference_suite_cuda(fname, cuda_api):
    ref_root = _reference_root()
    if ref_root is None:
        pytest.skip("the reference's tests are not staged under baseline/_ref (python baseline/stage_reference.py)")
    pytest.importorskip("pandas")
    _check(ref_root, fname, "product")
