"""The committed real-weight fixture (tests/golden/candy_head_golden.npz, generator tests/golden/make_candy_golden.py) as an
ONNX-style graph dict for shadernn_b200/onnx2snn.convert_graph: the first two stages of the reference's candy-9_simplified.onnx."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "candy_head_golden.npz")


def head_graph():
    z = np.load(GOLDEN)
    init = {k[5:]: z[k] for k in z.files if k.startswith("init/")}
    ops = [str(o) for o in z["node_ops"]]
    assert ops == ["Pad", "Conv", "InstanceNormalization", "Relu", "Pad", "Conv", "InstanceNormalization", "Relu"]
    names = sorted(init)  # conv1.conv2d.{bias,weight}, conv2.conv2d.{bias,weight}, in1.{bias,weight}, in2.{bias,weight}
    nodes, t = [], "input1"
    for s, (k, st, p) in enumerate([(9, 1, 4), (3, 2, 1)], start=1):
        nodes.append({"op": "Pad", "input": [t], "output": ["p%d" % s], "name": "pad%d" % s, "attr": {"mode": "reflect", "pads": [0, 0, p, p, 0, 0, p, p]}})
        nodes.append({"op": "Conv", "input": ["p%d" % s, "conv%d.conv2d.weight" % s, "conv%d.conv2d.bias" % s], "output": ["c%d" % s], "name": "conv%d" % s,
                      "attr": {"kernel_shape": [k, k], "strides": [st, st], "pads": [0, 0, 0, 0], "group": 1}})
        nodes.append({"op": "InstanceNormalization", "input": ["c%d" % s, "in%d.weight" % s, "in%d.bias" % s], "output": ["n%d" % s], "name": "in%d" % s,
                      "attr": {"epsilon": 1e-5}})
        nodes.append({"op": "Relu", "input": ["n%d" % s], "output": ["r%d" % s], "name": "relu%d" % s, "attr": {}})
        t = "r%d" % s
    assert all(n in names for nd in nodes for n in nd["input"][1:])
    return {"nodes": nodes, "init": init, "inputs": [("input1", [1, 3, 64, 64])], "outputs": [(t, [])]}, z["x"], z["y"]


