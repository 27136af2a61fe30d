"""SNNB_UMMA_TRACE=1 python tools/trace_1x1_probe.py 2> trace.txt : the MobileNetV2 1x1 layers that dominate its step, one launch each."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shadernn_b200 import core
ctx = core.GpuContext(0)
rng = np.random.default_rng(0)
for (n, hw, ic, oc, act) in [(64, 112, 16, 96, "relu6"), (64, 56, 144, 24, ""), (64, 56, 24, 144, "relu6"), (64, 14, 384, 64, "")]:
    x = rng.uniform(-1, 1, (n, hw, hw, ic)).astype(np.float32)
    w = (rng.standard_normal((oc, ic, 1, 1)) * 0.1).astype(np.float32)
    for i in range(2):
        core.conv2d(ctx, x, w, None, None, 1, 0, 0, "constant", act, 0.0, (hw, hw), algo="tcgen05")
