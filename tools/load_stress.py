"""Load + first forward + a few graph launches of one model, many times over in one process: flushes out launch-order races
(a hang traps after 2 s and surfaces as a load error). usage: load_stress.py [model] [repeats]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from shadernn_b200 import core, modelzoo

name = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
batch = {"resnet18": 32, "mobilenetv2": 64}.get(name, 8)
ctx = core.GpuContext(0)
d = tempfile.mkdtemp()
path, _ = modelzoo.build(name, d, input_hw=(224, 224))
x = modelzoo.synthetic_input(name, batch, (224, 224))
ref = {}
for i in range(reps):
    m = core.MixedInferenceCore(ctx, path, batch=batch, fuse=bool(i & 1), use_cuda_graph=bool(i & 1), precision="fp32x3")
    m.set_input(x)
    for _ in range(3):
        m.forward()
    out = m.get_output()
    assert np.array_equal(out, ref.setdefault(i & 1, out)), "load %d: output differs from the first load of this mode" % i
    del m
    print("load %d ok" % i, flush=True)
print("all %d loads ok" % reps)
