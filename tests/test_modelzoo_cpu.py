"""The synthetic BASELINE graphs (shadernn_b200/modelzoo.py) are worth comparing against:
activations stay O(1) through depth, the soft-max is not saturated, the arg-max differs from image to image — and the
oracle, walking the written JSON, agrees layer by layer with an independent torch-CPU evaluation of the same layer list
(modelzoo.torch_forward shares no code with the oracle or with the CUDA library), at BASELINE.json's input sizes."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import oracle  # noqa: E402
from shadernn_b200 import modelzoo  # noqa: E402


def _both(name, hw, batch, tmp, **kw):
    path, layers = modelzoo.build(name, tmp, input_hw=hw, **kw)
    x = modelzoo.synthetic_input(name, batch, hw)
    want = oracle.Model(path).run(x, return_all=True)
    last = len(layers) - 1 - (1 if layers[-1]["type"] in ("YOLO", "Lambda") else 0)
    got = modelzoo.torch_forward(layers, x, upto=last)
    return layers, want, got, last


@pytest.mark.parametrize("name,hw,batch", [("resnet18", (224, 224), 8), ("mobilenetv2", (224, 224), 4), ("yolov3tiny", (416, 416), 1), ("candy", (144, 144), 1),
                                           ("espcn", (224, 224), 1)])
def test_oracle_agrees_with_torch_on_every_layer(tmp_path, name, hw, batch):
    layers, want, got, last = _both(name, hw, batch, str(tmp_path))
    for i in range(last + 1):
        assert want[i].shape == got[i].shape, (i, layers[i]["type"], want[i].shape, got[i].shape)
        scale = float(np.abs(want[i]).max())
        err = float(np.abs(want[i] - got[i]).max())
        assert err <= 5e-5 * max(scale, 1e-6), "layer %d (%s): oracle vs torch %.3g of range %.3g" % (i, layers[i]["type"], err, scale)
        if layers[i]["type"] in ("Conv2D", "DepthwiseConv2D", "Add") and name != "candy":
            rms = float(np.sqrt(np.mean(want[i].astype(np.float64) ** 2)))
            assert 0.05 < rms < 8.0, "layer %d (%s): activations left O(1): rms %.3g" % (i, layers[i]["type"], rms)


@pytest.mark.parametrize("name,classes,batch", [("resnet18", 10, 16), ("mobilenetv2", 1000, 8)])
def test_classifier_heads_discriminate_between_images(tmp_path, name, classes, batch):
    path, layers = modelzoo.build(name, str(tmp_path), input_hw=(224, 224))
    x = modelzoo.synthetic_input(name, batch, (224, 224))
    probs = oracle.Model(path).run(x).reshape(batch, classes)
    assert np.allclose(probs.sum(1), 1.0, atol=1e-4)
    assert float(probs.max()) < 0.999, "soft-max saturated"
    assert len(set(probs.argmax(1).tolist())) >= 4, probs.argmax(1)
    # the pre-soft-max variant used by the logit-level parity tests
    path2, _ = modelzoo.build(name, str(tmp_path / "lin"), input_hw=(224, 224), head_activation="linear")
    logits = oracle.Model(path2).run(x).reshape(batch, classes)
    assert 0.3 < float(logits.std()) < 20.0
    assert np.array_equal(logits.argmax(1), probs.argmax(1))
