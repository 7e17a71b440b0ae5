"""Runs the REFERENCE'S OWN golden-vector tests (tests/test_curated_examples.py +
tests/assets/curated_examples/*.json of /root/reference) against this package: `cmtj` is aliased
to cmtj_b200 (install_as_cmtj), the reference's pure-Python `cmtj.utils` helpers are picked up
from the reference tree on top of it, and the compute goes through our facade, packing code and
the oracle restatement.  This is the "golden vectors of the reference" pin of SURVEY 8(c) for
the CPU side; the same scripts run on the GPU path in tests/test_gpu_reference_suite.py.

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
