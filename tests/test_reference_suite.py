"""Runs the REFERENCE'S OWN golden-vector tests (tests/test_curated_examples.py +
tests/assets/curated_examples/*.json of /root/reference) against this package: `cmtj` is aliased
to cmtj_b200 (install_as_cmtj), the reference's pure-Python `cmtj.utils` helpers are picked up
from the reference tree on top of it, and the compute goes through our facade, packing code and
the oracle restatement.  This is the "golden vectors of the reference" pin of SURVEY 8(c) for
the CPU side.  The second half runs the reference's OTHER test files for the path (drivers, stack,
group interaction, solvers, SOT torques, constants, sanity, basic, procedures) unmodified, in a
subprocess, with `cmtj` aliased to this package (tests/ref_alias_plugin.py).

CPU only, and only where /root/reference exists (it does not travel to the GPU box)."""
import importlib.util
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present")

FAST = ["test_sto", "test_harmonic_hall", "test_cims_stability", "test_fmr_basic", "test_ahe_loops", "test_cims_hysteresis",
        "test_vsd_freq_scan", "test_vsd_field_scan"]


def load_reference_tests():
    import cmtj_b200

    cmtj_b200.install_as_cmtj()
    ref_pkg = os.path.join(REF, "cmtj")
    if ref_pkg not in cmtj_b200.__path__:
        cmtj_b200.__path__.append(ref_pkg)  # cmtj.models etc.: the reference's pure-Python helpers
    # `cmtj.utils` must be the REFERENCE's numpy/scipy helpers here (ours reduce on the GPU and this is
    # the CPU-side pin of the facade + oracle); install_as_cmtj pointed it at cmtj_b200.utils.
    for k in [k for k in sys.modules if k == "cmtj.utils" or k.startswith("cmtj.utils.")]:
        del sys.modules[k]
    uspec = importlib.util.spec_from_file_location("cmtj.utils", os.path.join(ref_pkg, "utils", "__init__.py"),
                                                   submodule_search_locations=[os.path.join(ref_pkg, "utils")])
    umod = importlib.util.module_from_spec(uspec)
    sys.modules["cmtj.utils"] = umod
    uspec.loader.exec_module(umod)
    name = "_ref_curated_examples"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "tests", "test_curated_examples.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", FAST)
def test_reference_golden_vectors_through_our_facade(on_oracle, name):
    import cmtj_b200

    mod = load_reference_tests()
    assert mod.Junction is cmtj_b200.Junction and mod.cmtj is cmtj_b200  # really our facade, not the reference's
    getattr(mod, name)(False)  # update_reference=False: compare with the reference's committed JSON


REF_TEST_FILES = ["test_basic", "test_drivers", "test_sanity", "test_stack", "test_group_interaction", "test_constants",
                  "test_sot_torques", "test_solvers"]


def _run_reference_files(names, tmp_path):
    import subprocess

    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([here, os.path.dirname(here), os.environ.get("PYTHONPATH", "")]))
    cmd = [sys.executable, "-m", "pytest"] + [os.path.join(REF, "tests", n + ".py") for n in names] + [
        "-p", "ref_alias_plugin", "-p", "no:cacheprovider", "-q", "--rootdir", str(tmp_path), "-c", os.devnull]
    r = subprocess.run(cmd, env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout and "failed" not in r.stdout, r.stdout[-500:]
    return r.stdout


def test_reference_test_files_pass_against_this_package(tmp_path):
    """`pytest /root/reference/tests/{test_basic,test_drivers,...}.py` with `import cmtj` resolving to cmtj_b200
    (facade on the oracle): the reference's own assertions on API surface, error messages and numbers.
    One subprocess for the eight files (42 tests)."""
    out = _run_reference_files(REF_TEST_FILES, tmp_path)
    n = int(out.strip().splitlines()[-1].split(" passed")[0].split()[-1])
    assert n >= 42, out[-300:]


def test_reference_procedures_test_file_passes_against_this_package(tmp_path):
    """tests/test_procedures.py of the reference (PIMM / VSD procedures of cmtj.utils driving OUR Junction)."""
    _run_reference_files(["test_procedures"], tmp_path)
