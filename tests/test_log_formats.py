"""On-disk log formats: Junction.saveLog / Stack.saveLogs write what the reference's
saveLogs writes (junction.hpp:1638-1656, stack.hpp:222-241) -- same keys, same `%g` text per value;
the column ORDER of the reference is its unordered_map hash order and is not part of the format --
and the batched sweep log round-trips through .npz losslessly."""
import ctypes
import ctypes.util

import numpy as np
import pytest

import cmtj_b200 as cm
import helpers

libc = ctypes.CDLL(ctypes.util.find_library("c"))


def _read_csv(path):
    rows = [line.rstrip("\n") for line in open(path)]
    assert all(r.endswith(";") for r in rows)
    keys = rows[0].split(";")[:-1]
    cols = {k: [] for k in keys}
    for r in rows[1:]:
        vals = r.split(";")[:-1]
        assert len(vals) == len(keys)
        for k, v in zip(keys, vals):
            cols[k].append(v)
    return cols


def _run(mod):
    j = helpers.two_layer_mtj(mod)
    j.setLayerExternalFieldDriver("all", mod.AxialDriver(mod.constantDriver(120e3), mod.constantDriver(-3e3), mod.constantDriver(8e3)))
    j.runSimulation(2e-11, 1e-13, 1e-12)
    return j


def _c_g(x):
    """What `ostream << double` writes with default flags: printf("%g") (C library, not Python's formatter)."""
    buf = ctypes.create_string_buffer(64)
    libc.snprintf(buf, 64, b"%g", ctypes.c_double(x))
    return buf.value.decode()


def test_savelog_is_the_references_stream_format(on_oracle, tmp_path):
    """The reference's own saveLog cannot serve as the checker here: the unmodified extension
    segfaults inside Junction::saveLogs in this container (observed for every fixture), so the
    format is pinned from its source (junction.hpp:1638-1656): header `key;` per key, then one row
    per sample of `value;` in the stream default."""
    j = _run(cm)
    a = tmp_path / "ours.csv"
    j.saveLog(str(a))
    cols, log = _read_csv(a), j.getLog()
    assert set(cols) == set(log)
    for k, v in log.items():
        assert cols[k] == [_c_g(x) for x in v], k


def test_savelog_rejects_an_empty_filename(on_oracle):
    with pytest.raises(RuntimeError, match="may not be empty"):
        _run(cm).saveLog("")


def test_stack_savelogs_format(on_oracle, tmp_path, capsys):
    js = [helpers.two_layer_mtj(cm) for _ in range(2)]
    st = cm.stack.SeriesStack(js, "free", "bottom", 0.0, True)
    st.setCoupledCurrentDriver(cm.constantDriver(1e9))
    st.setCouplingStrength(0.1)
    st.runSimulation(1e-11, 1e-13, 1e-12)
    a = tmp_path / "stack.csv"
    st.saveLogs(str(a))
    cols, log = _read_csv(a), st.getLog()
    assert set(cols) == set(log)
    for k, v in log.items():
        assert cols[k] == [_c_g(x) for x in v], k
    st.saveLogs("")  # stack.hpp:223-227: an empty filename is ignored with a message, not an error
    assert "Ignoring file save" in capsys.readouterr().out


def test_batch_log_npz_round_trip_and_csv(on_oracle, tmp_path):
    j = helpers.two_layer_mtj(cm)
    b = cm.JunctionBatch(j, 5)
    b.set_external_field("all", np.linspace(-1e5, 1e5, 15).reshape(5, 3))
    log = b.run(2e-11, 1e-13, 1e-12)
    p = tmp_path / "sweep.npz"
    log.save_npz(str(p))
    back = cm.BatchLog.load_npz(str(p))
    assert set(back) == set(log)
    for k in log:
        assert np.array_equal(back[k], log[k])
    # instance 3 as CSV == what a single Junction with that field writes
    j1 = helpers.two_layer_mtj(cm)
    H = np.linspace(-1e5, 1e5, 15).reshape(5, 3)[3]
    j1.setLayerExternalFieldDriver("all", cm.AxialDriver(*[cm.constantDriver(float(h)) for h in H]))
    j1.runSimulation(2e-11, 1e-13, 1e-12)
    j1.saveLog(str(tmp_path / "one.csv"))
    log.save_csv(str(tmp_path / "inst.csv"), 3)
    one, inst = _read_csv(tmp_path / "one.csv"), _read_csv(tmp_path / "inst.csv")
    assert set(one) == set(inst)
    for k in one:
        assert one[k] == inst[k], k
