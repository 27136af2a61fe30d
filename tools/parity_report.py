#!/usr/bin/env python
"""Per-layer parity report of one BASELINE workload on the GPU: the engine runs the full batch, the oracle a sample of it, and
every observable layer's max|err| / max|reference| is printed (the north_star's "1e-3 relative per layer" criterion).
  python tools/parity_report.py resnet18 --batch 32 --sample 4 --precision fp16w [--fuse]
TEST INFRASTRUCTURE (imports oracle/); never part of the product path."""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402
from shadernn_b200 import core, modelzoo  # noqa: E402
from shadernn_b200._lib import SnnbError  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--sample", type=int, default=2)
    ap.add_argument("--hw", type=int, default=0)
    ap.add_argument("--precision", default="fp32x3")
    ap.add_argument("--fuse", action="store_true")
    a = ap.parse_args()
    hw = (a.hw, a.hw) if a.hw else modelzoo.MODELS[a.model][1]
    d = tempfile.mkdtemp(prefix="snnb_parity_")
    path, layers = modelzoo.build(a.model, d, input_hw=hw)
    x = modelzoo.synthetic_input(a.model, a.batch, hw)
    want = oracle.Model(path).run(x[:a.sample], return_all=True)
    ctx = core.GpuContext(0)
    m = core.MixedInferenceCore(ctx, path, batch=a.batch, input_hw=hw, fuse=a.fuse, precision=a.precision)
    m.set_input(x)
    m.forward()
    ctx.sync()
    print("# %s %dx%d batch %d (oracle on %d images) precision %s fuse %d: max|err| / max|ref| per layer" % (a.model, hw[0], hw[1], a.batch, a.sample, a.precision, a.fuse))
    worst = 0.0
    for i in range(m.num_layers):
        name, typ, shape = m.layer_info(i)
        if typ == "YOLO":
            continue
        try:
            got = m.layer_output(i)[:a.sample]
        except SnnbError:
            print("[%02d] %-22s fused away" % (i, typ))
            continue
        if got.shape != want[i].shape or (a.fuse and layers[i]["type"] in ("ZeroPadding2D", "Flatten")):
            print("[%02d] %-22s alias (fused)" % (i, typ))
            continue
        scale = float(np.abs(want[i]).max())
        err = float(np.abs(got.astype(np.float64) - want[i]).max())
        worst = max(worst, err / max(scale, 1e-30))
        print("[%02d] %-22s %-18s range %9.4g  max|err| %9.3g  rel %8.2e" % (i, typ, "x".join(map(str, shape[1:])), scale, err, err / max(scale, 1e-30)))
    print("# worst %.3g (limit 1e-3)" % worst)


if __name__ == "__main__":
    main()
