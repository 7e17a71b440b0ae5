"""Achieved-parity ledger: every parity test records what it measured -- the achieved error (or z value), the
bound applied and whether a sensitivity widening was needed -- and the session writes the ledger to
``gpurun_out/parity_<tag>.json`` (copied to profiles/ by tools/gpu_round.sh).  Test infrastructure only."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENTRIES = []


def record(what, **fields):
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
    e = {"test": test, "what": what}
    for k, v in fields.items():
        e[k] = v if isinstance(v, (str, bool, int)) or v is None else float(v)
    ENTRIES.append(e)


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
