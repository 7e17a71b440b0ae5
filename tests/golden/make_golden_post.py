#!/usr/bin/env python
"""Generates tests/golden/post_*.npz by running tests/golden/post_scenarios.py against the
UNMODIFIED reference: the oracle/_ref extension aliased as `cmtj`, with the reference's own pure-Python
cmtj/utils/procedures.py imported from /root/reference on top of it.
    python tests/golden/make_golden_post.py [--out DIR]
Needs /root/reference (build container only); the .npz files are committed."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import helpers  # noqa: E402

REF = "/root/reference"


def reference_procedures():
    ref = helpers.reference_module()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    sys.modules["cmtj"] = ref
    ref.__path__ = [os.path.join(REF, "cmtj")]
    for sub in ("constants", "stack", "reservoir", "noise", "llgb"):
        if hasattr(ref, sub):
            sys.modules["cmtj." + sub] = getattr(ref, sub)
    import cmtj.utils.procedures as P  # the reference's file, unmodified

    assert P.__file__.startswith(REF)
    return ref, P


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    ref, P = reference_procedures()
    import post_scenarios

    for name, fn in post_scenarios.SCENARIOS.items():
        if a.only and name != a.only:
            continue
        out = fn(ref, P, False)
        path = os.path.join(a.out, "post_" + name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(out)} arrays, {os.path.getsize(path)} bytes")
