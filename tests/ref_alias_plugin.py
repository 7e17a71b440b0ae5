"""pytest plugin (-p ref_alias_plugin): run the REFERENCE's own test files against this package.  `cmtj` is aliased to
cmtj_b200 before the reference's conftest.py is imported, and the facade's backend is the CPU oracle (test
infrastructure: only tests may route the facade there).  Used by tests/test_reference_suite.py."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
for p in (HERE, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

import cmtj_b200  # noqa: E402
import helpers  # noqa: E402
from cmtj_b200 import _abi  # noqa: E402

cmtj_b200.install_as_cmtj()
_oracle = helpers.oracle_backend()
_abi.backend = lambda: _oracle

# the reference's pure-Python subpackages (cmtj.models, cmtj.utils: numpy / scipy / sympy helpers outside the
# hot path) come from the reference tree; everything that integrates is ours
import importlib.util  # noqa: E402

REF_PKG = "/root/reference/cmtj"
if REF_PKG not in cmtj_b200.__path__:
    cmtj_b200.__path__.append(REF_PKG)
for k in [k for k in sys.modules if k == "cmtj.utils" or k.startswith("cmtj.utils.")]:
    del sys.modules[k]
_uspec = importlib.util.spec_from_file_location("cmtj.utils", os.path.join(REF_PKG, "utils", "__init__.py"),
                                                submodule_search_locations=[os.path.join(REF_PKG, "utils")])
_umod = importlib.util.module_from_spec(_uspec)
sys.modules["cmtj.utils"] = _umod
_uspec.loader.exec_module(_umod)
