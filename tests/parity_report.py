"""Achieved-parity ledger: every parity test records what it measured -- the achieved error (or z value), the
bound applied and whether a sensitivity widening was needed -- and the session writes the ledger to
``gpurun_out/parity_<tag>.json`` (copied to profiles/ by tools/gpu_round.sh).  Test infrastructure only."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENTRIES = []


def record(what, **fields):
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
    e = {"test": test, "what": what}
    for k, v in fields.items():
        e[k] = v if isinstance(v, (str, bool, int, list)) or v is None else float(v)
    ENTRIES.append(e)
    return e


SEPARATED = 1e-3  # a 1e-12 perturbation grown to this size: the trajectory has left the linear regime


def compare_with_sensitivity(what, keys, got, want, perts):
    """per instance and channel: relative error (scale = the channel's own max over time) against
    max(1e-9, 10 x the reference's own sensitivity up to that sample), the sensitivity being the largest response of
    the ORACLE to a 1e-12 change of one component of m0 (``perts``: one perturbed oracle run per component, all
    channels pooled per instance).  Once that response has grown to SEPARATED the reference trajectory itself is
    not reproducible (any two implementations, or two compilers, part ways there) and later samples of that instance
    are not compared.  Records the achieved error and how many instances needed the widened bound."""
    if isinstance(perts, dict):
        perts = [perts]

    def per_instance(a):  # [T, I, ...] -> [T, I]: an instance is everything behind axis 1 (the junctions of a stack)
        return a.reshape(a.shape[0], a.shape[1], -1).max(axis=2)

    scale = {k: np.maximum(np.max(np.abs(want[k]), axis=0, keepdims=True), 1e-300) for k in keys}
    sens = 0.0
    for k in keys:  # pooled over channels: an instability shows up in every channel, not at the same sample in each
        for pert in perts:
            sens = np.maximum(sens, per_instance(np.abs(pert[k] - want[k]) / scale[k]))
    sens = np.maximum.accumulate(sens, axis=0)
    bound = np.maximum(1e-9, 10 * sens)
    live = sens < SEPARATED
    worst, worst_live, n_wide = 0.0, 0.0, np.zeros(sens.shape[-1], dtype=bool)
    err_inst = np.zeros(sens.shape[-1])
    for k in keys:
        err = per_instance(np.abs(got[k] - want[k]) / scale[k])
        worst = max(worst, float(err.max()))
        worst_live = max(worst_live, float(err[live].max()))
        err_inst = np.maximum(err_inst, err.max(axis=0))
        n_wide |= (err > 1e-9).any(axis=0)
        bad = (err > bound) & live
        assert not bad.any(), (what, k, float(err[bad].max()), float(sens[bad].min()), int(bad.any(axis=0).sum()))
    return record(what, max_rel_err=worst, bound_1e9=1e-9, max_bound_applied=float(bound[live].max()), widened=bool(n_wide.any()),
           instances=int(n_wide.size), instances_over_1e9=int(n_wide.sum()),
           instances_separated=int((~live).any(axis=0).sum()), max_rel_err_before_separation=worst_live,
           max_rel_err_of_the_others=float(err_inst[~n_wide].max()) if (~n_wide).any() else None,
           err_per_instance=[float(v) for v in err_inst] if err_inst.size <= 64 else None,
           reference_sensitivity_per_instance=[float(v) for v in sens[-1]] if err_inst.size <= 64 else None)


def write(tag="r02"):
    if not ENTRIES:
        return None
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"parity_{tag}.json")
    # several pytest sessions of one GPU visit fill one ledger: entries of earlier sessions are kept unless this
    # session measured the same (test, what) again; CMTJB_PARITY_FRESH=1 starts over
    old = []
    if os.path.exists(path) and not os.environ.get("CMTJB_PARITY_FRESH"):
        try:
            mine = {(e["test"], e["what"]) for e in ENTRIES}
            old = [e for e in json.load(open(path))["entries"] if (e["test"], e["what"]) not in mine]
        except Exception:
            old = []
    det = [e for e in old + ENTRIES if "max_rel_err" in e]
    summary = {
        "deterministic_entries": len(det),
        "worst_max_rel_err": max((e["max_rel_err"] for e in det), default=None),
        "entries_over_1e-9": sum(1 for e in det if e["max_rel_err"] > 1e-9),
        "entries_needing_widened_bound": sum(1 for e in det if e.get("widened")),
        "stochastic_entries": sum(1 for e in old + ENTRIES if "z" in e),
        "worst_abs_z": max((abs(e["z"]) for e in old + ENTRIES if "z" in e), default=None),
    }
    json.dump({"summary": summary, "entries": old + ENTRIES}, open(path, "w"), indent=1)
    return path
