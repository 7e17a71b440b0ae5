"""The built libcmtj_b200.so loads without a GPU and exports every function that
include/cmtj_b200.h declares; the ctypes mirrors have the header's struct sizes."""
import ctypes as C
import os
import re
import subprocess

import pytest

from cmtj_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cmtj_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmtjb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("cmtjb_plan", "cmtjb_run_host", "cmtjb_run_device", "cmtjb_stack_run_host", "cmtjb_group_run_host",
                 "cmtjb_last_error", "cmtjb_abi_version", "cmtjb_run_device_post", "cmtjb_strip_resistance_device"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()  # dlopen only; no CUDA call
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.cmtjb_abi_version() == _abi.ABI_VERSION


def test_struct_sizes_match_the_header(tmp_path):
    prog = tmp_path / "sizes.c"
    prog.write_text(
        '#include <stdio.h>\n#include "cmtj_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
        "sizeof(cmtjb_driver_t),sizeof(cmtjb_layer_t),sizeof(cmtjb_junction_t),sizeof(cmtjb_override_t),"
        "sizeof(cmtjb_run_args_t),sizeof(cmtjb_plan_t),sizeof(cmtjb_stack_t),sizeof(cmtjb_group_t),sizeof(cmtjb_coupled_plan_t),sizeof(cmtjb_post_t));return 0;}\n"
    )
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(t) for t in (_abi.Driver, _abi.Layer, _abi.Junction, _abi.Override, _abi.RunArgs, _abi.Plan,
                                  _abi.Stack, _abi.Group, _abi.CoupledPlan, _abi.Post)]
    assert got == want


def test_no_gpu_means_loud_failure_not_fallback():
    """On a machine without a CUDA device every compute call raises; nothing routes to the CPU."""
    lib = _abi.load_library()
    if lib.cmtjb_device_count() > 0:
        pytest.skip("a GPU is visible")
    import cmtj_b200
    import helpers

    j = helpers.two_layer_mtj(cmtj_b200)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        j.runSimulation(1e-10, 1e-13, 1e-13)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        cmtj_b200.JunctionBatch(j, 4).run(1e-10, 1e-13, 1e-13)
