import sys, numpy as np
sys.path.insert(0,'/root/repo')
from shadernn_b200 import core
ctx = core.GpuContext(0)
rng = np.random.default_rng(0)
x = rng.uniform(-1,1,(32,56,56,64)).astype(np.float32)
w = (rng.standard_normal((64,64,3,3))*0.05).astype(np.float32)
for i in range(2):
    core.conv2d(ctx, x, w, None, None, 1, 1, 1, "constant", "relu", 0.0, (56,56), algo="tcgen05")
x = rng.uniform(-1,1,(32,28,28,128)).astype(np.float32)
w = (rng.standard_normal((128,128,3,3))*0.05).astype(np.float32)
for i in range(2):
    core.conv2d(ctx, x, w, None, None, 1, 1, 1, "constant", "relu", 0.0, (28,28), algo="tcgen05")
