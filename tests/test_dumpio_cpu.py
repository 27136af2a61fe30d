"""Per-layer dump tooling (shadernn_b200/dumpio.py, tools/compare_dumps.py; SURVEY §8 f-N2): the reference's `.dump` format
(core/src/image.cpp:216-245) read, written and compared on the host. The GPU suite checks that `snnb_tensor_dump` emits the same
bytes (tests/test_ops_gpu.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shadernn_b200 import dumpio  # noqa: E402


@pytest.mark.parametrize("shape", [(5, 7, 6), (1, 1, 10), (3, 4, 4), (2, 3, 1)])
def test_round_trip_and_layout(tmp_path, shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.uniform(-5, 5, shape).astype(np.float32)
    p = str(tmp_path / "layer.dump")
    dumpio.write_dump(p, x)
    raw = open(p, "rb").read()
    h, w, c = shape
    d = (c + 3) // 4
    assert raw[:32].rstrip(b"\0") == ("%d %d %d %d" % (w, h, d, c)).encode()
    assert len(raw) == 32 + d * h * w * 4 * 4
    c4 = np.frombuffer(raw[32:], np.float32).reshape(d, h, w, 4)
    assert c4[0, 0, 0, 0] == x[0, 0, 0]                       # slice 0 holds channels 0..3
    assert np.all(c4[d - 1, :, :, (c - 1) % 4 + 1:] == 0)      # padding lanes of the last slice are zero
    assert np.array_equal(dumpio.read_dump(p), x)


def test_comparator_is_abs_and_rel():
    a = np.array([100.0, 0.001, 1.0], np.float32)
    assert dumpio.mismatches(a, a)[0] == 0
    assert dumpio.mismatches(a, a + np.array([0.5, 0, 0], np.float32))[0] == 0     # large abs, tiny rel
    assert dumpio.mismatches(a, a + np.array([0, 0.005, 0], np.float32))[0] == 0   # large rel, tiny abs
    assert dumpio.mismatches(a, a + np.array([0, 0, 0.5], np.float32))[0] == 1


def test_compare_dirs_cli(tmp_path):
    da, db = tmp_path / "a", tmp_path / "b"
    da.mkdir(), db.mkdir()
    rng = np.random.default_rng(1)
    for i, shape in enumerate([(4, 4, 8), (2, 2, 3)]):
        x = rng.uniform(-1, 1, shape).astype(np.float32)
        dumpio.write_dump(str(da / ("m layer [%02d] Conv2D.dump" % i)), x)
        dumpio.write_dump(str(db / ("m layer [%02d] Conv2D.dump" % i)), x + (0.5 if i == 1 else 1e-4))
    dumpio.write_dump(str(da / "m layer [07] Add.dump"), np.zeros((1, 1, 4), np.float32))
    rows = dumpio.compare_dirs(str(da), str(db))
    assert [r[4] for r in rows][:2] == ["ok", "MISMATCH"] and rows[2][4].startswith("only in")
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "compare_dumps.py")
    r = subprocess.run([sys.executable, tool, str(da), str(db)], capture_output=True, text=True)
    assert r.returncode == 1 and "MISMATCH" in r.stdout and "3 file(s), 2 not ok" in r.stdout


def test_bad_file_is_rejected(tmp_path):
    p = tmp_path / "x.dump"
    p.write_bytes(b"not a dump".ljust(40, b"\0"))
    with pytest.raises(ValueError):
        dumpio.read_dump(str(p))
