"""Time the ResNet-18 stem (7x7 s2, 224x224x3 -> 112x112x64, batch 32) alone; with SNNB_UMMA_ABLATE=2 the activation loads are skipped
(results wrong): how much of the kernel is the TMA element rate?"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from shadernn_b200 import core
from shadernn_b200._lib import lib, check
ctx = core.GpuContext(0)
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, (32, 224, 224, 3)).astype(np.float32)
w = (rng.standard_normal((64, 3, 7, 7)) * 0.1).astype(np.float32)
d = core.conv_desc(3, 64, 7, 2, 3, 3, "constant", "relu", 0.0, "tcgen05")
wh = core.vp()
check(lib().snnb_weights_pack_conv2d(ctx.h, C.byref(d), w.ctypes.data_as(core.vp), None, None, None, None, None, C.byref(wh)))
tin = core.ImageTexture.from_numpy(ctx, x)
tout = core.ImageTexture(ctx, 32, 112, 112, 64)
for _ in range(5):
    check(lib().snnb_conv2d_launch(ctx.h, C.byref(d), wh, tin.h, None, tout.h))
ctx.sync()
tm = C.c_void_p()
check(lib().snnb_timer_create(ctx.h, C.byref(tm)))
check(lib().snnb_timer_start(tm))
for _ in range(50):
    check(lib().snnb_conv2d_launch(ctx.h, C.byref(d), wh, tin.h, None, tout.h))
check(lib().snnb_timer_stop(tm))
ms = C.c_float()
check(lib().snnb_timer_elapsed_ms(tm, C.byref(ms)))
print("stem 7x7 s2 batch 32: %.1f us per launch (ablate=%s)" % (ms.value / 50 * 1e3, os.environ.get("SNNB_UMMA_ABLATE", "0")))
