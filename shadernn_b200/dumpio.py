"""Reader / writer / comparator for ShaderNN per-layer `.dump` files (SURVEY §8 f-N2).

Format (core/src/image.cpp:216-245, read by the reference's tools/misc/readTextureDump.py): a 32-byte NUL-padded ASCII header
"W H D C" (width, height, depth = ceil(C/4) RGBA slices, channels) followed by D*H*W*4 little-endian float32 in C4HW4 order
[slice][y][x][rgba]. `snnb_model_dump_outputs` (--dump_outputs, vulkanBackend.cpp:108-143) writes one file per layer under the
reference's layer names ("<model> layer [NN] <Type>.dump", batch index appended as ".nK" when N > 1), so a directory written here
can be diffed against one written by a real ShaderNN run on another machine - which is how the "parity vs ncnn unpinned" gap
closes for someone who has the LFS model files.
"""
import os

import numpy as np

HEADER_BYTES = 32


def write_dump(path, hwc):
    """Write one image (H, W, C) float32 as a reference-format dump file."""
    a = np.ascontiguousarray(hwc, dtype=np.float32)
    h, w, c = a.shape
    d = (c + 3) // 4
    c4 = np.zeros((d, h, w, 4), np.float32)
    for s in range(d):
        n = min(4, c - 4 * s)
        c4[s, :, :, :n] = a[:, :, 4 * s:4 * s + n]
    hdr = ("%d %d %d %d" % (w, h, d, c)).encode("ascii")
    if len(hdr) >= HEADER_BYTES:
        raise ValueError("dump header too long: %r" % hdr)
    with open(path, "wb") as f:
        f.write(hdr.ljust(HEADER_BYTES, b"\0"))
        f.write(c4.tobytes())


def read_dump(path):
    """Read a dump file back as (H, W, C) float32."""
    raw = open(path, "rb").read()
    fields = raw[:HEADER_BYTES].rstrip(b"\0").split()
    if len(fields) != 4:
        raise ValueError("%s: not a ShaderNN dump (header %r)" % (path, raw[:HEADER_BYTES]))
    w, h, d, c = (int(x) for x in fields)
    body = np.frombuffer(raw[HEADER_BYTES:], dtype="<f4")
    if body.size != d * h * w * 4 or d != (c + 3) // 4:
        raise ValueError("%s: payload of %d floats does not match header W=%d H=%d D=%d C=%d" % (path, body.size, w, h, d, c))
    c4 = body.reshape(d, h, w, 4)
    return np.ascontiguousarray(np.moveaxis(c4, 0, 2).reshape(h, w, d * 4)[:, :, :c])


def mismatches(a, b, eps=0.01):
    """The reference's comparator (demo/common/testutil.cpp:351-361): an element differs when BOTH its absolute and its relative
    error exceed eps. Returns (count, max abs error)."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    if a.shape != b.shape:
        return a.size, float("inf")
    diff = np.abs(a - b)
    rel = diff / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-30)
    bad = (diff > eps) & (rel > eps)
    return int(bad.sum()), float(diff.max()) if diff.size else 0.0


def compare_dirs(dir_a, dir_b, eps=0.01):
    """Compare every dump file present in both directories; returns rows (name, shape, mismatches, max_abs_err, status)."""
    rows = []
    names_a = {f for f in os.listdir(dir_a) if ".dump" in f}
    names_b = {f for f in os.listdir(dir_b) if ".dump" in f}
    for name in sorted(names_a | names_b):
        if name not in names_a or name not in names_b:
            rows.append((name, None, None, None, "only in " + (dir_a if name in names_a else dir_b)))
            continue
        a, b = read_dump(os.path.join(dir_a, name)), read_dump(os.path.join(dir_b, name))
        if a.shape != b.shape:
            rows.append((name, a.shape, None, None, "shape %s vs %s" % (a.shape, b.shape)))
            continue
        bad, mx = mismatches(a, b, eps)
        rows.append((name, a.shape, bad, mx, "ok" if bad == 0 else "MISMATCH"))
    return rows
