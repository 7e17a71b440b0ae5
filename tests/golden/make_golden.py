#!/usr/bin/env python
"""Generates tests/golden/*.npz by running tests/golden/scenarios.py against the UNMODIFIED
reference extension (oracle/_ref/_cmtj*.so, built from /root/reference by oracle/Makefile).
Run in the build container:  python tests/golden/make_golden.py
Also copies the reference's own direct-trajectory golden vector (tests/assets/curated_examples/
test_sto.json, 400 numbers rounded to 3 decimals) next to them as ref_test_sto.json."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import helpers  # noqa: E402
import scenarios  # noqa: E402

ref = helpers.reference_module()
assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
for name, fn in scenarios.SCENARIOS.items():
    out = fn(ref)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(out)} arrays, {os.path.getsize(path)} bytes")
src = "/root/reference/tests/assets/curated_examples/test_sto.json"
if os.path.exists(src):
    json.dump(json.load(open(src)), open(os.path.join(HERE, "ref_test_sto.json"), "w"))
