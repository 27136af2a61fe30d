import sys, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import oracle
from shadernn_b200 import core
ctx = core.GpuContext(0)
rng = np.random.default_rng(0)
for amp in (1.0, 0.05, 0.001):
    x = (rng.uniform(-1,1,(2,28,28,128))*amp).astype(np.float32)
    w = (rng.standard_normal((128,128,3,3))*np.sqrt(2/(9*128))).astype(np.float32)
    want = oracle.conv2d(x, w, None, None, 1, 1, 1, "constant", "", 0.0, (28,28))
    for prec in ("fp32x3","fp16w"):
        ctx.set_precision(prec)
        got = core.conv2d(ctx, x, w, None, None, 1, 1, 1, "constant", "", 0.0, (28,28), algo="tcgen05")
        simt = core.conv2d(ctx, x, w, None, None, 1, 1, 1, "constant", "", 0.0, (28,28), algo="simt")
        s = float(np.abs(want).max())
        print("amp %g %s: tcgen05 rel err %.3g   simt rel err %.3g" % (amp, prec, float(np.abs(got-want).max())/s, float(np.abs(simt-want).max())/s))
ctx.set_precision("fp32x3")
