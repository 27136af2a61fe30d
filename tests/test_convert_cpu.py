"""torch / torchvision -> ShaderNN JSON writer (shadernn_b200/convert.py, SURVEY §8 f-N1), checked end to end on the CPU:
the exported file, read back by the oracle's independent reader and run with the oracle's operators, must reproduce what
torch itself computes for the same module and input (torch-CPU is the secondary cross-check of SURVEY §8c: it shares no
code with either the oracle or the CUDA library)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

torch = pytest.importorskip("torch")
torchvision = pytest.importorskip("torchvision")

from shadernn_b200 import convert  # noqa: E402
from oracle import oracle  # noqa: E402


def _randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.uniform_(0.5, 1.5, generator=g)
            m.bias.data.uniform_(-0.1, 0.1, generator=g)
            m.running_mean.uniform_(-0.1, 0.1, generator=g)
            m.running_var.uniform_(0.5, 1.5, generator=g)


def make(name, classes, seed=7767517):
    torch.manual_seed(seed)
    model = getattr(torchvision.models, name)(weights=None, num_classes=classes)
    _randomise_bn(model, seed + 1)
    return model.eval()


@pytest.mark.parametrize("name,hw,split", [("resnet18", (64, 64), False), ("resnet18", (96, 96), True), ("mobilenet_v2", (64, 64), True)])
def test_exported_model_matches_torch(tmp_path, name, hw, split):
    model = make(name, classes=20)
    path = str(tmp_path / (name + ".json"))
    layers = convert.export(model, path, input_hw=hw, split=split)
    assert layers[0]["type"] == "InputLayer" and layers[-1]["type"] == "Dense"
    x = np.random.default_rng(3).uniform(-1, 1, (2, hw[0], hw[1], 3)).astype(np.float32)
    with torch.no_grad():
        want = model(torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()).numpy()
    got = oracle.Model(path).run(x).reshape(2, -1)
    scale = float(np.abs(want).max())
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) <= 1e-3 * scale, (float(np.abs(got - want).max()), scale)
    assert np.array_equal(got.argmax(1), want.argmax(1))


def test_bn_eps_is_folded_into_the_variance():
    # the reference hard-codes eps = 1e-3 (vk_conv2d.comp:282-283): a module with eps 1e-5 must be written so that
    # var_json + 1e-3 == var + 1e-5
    bn = torch.nn.BatchNorm2d(4, eps=1e-5)
    bn.running_var.fill_(0.75)
    p = convert.TorchExporter._bn_params(bn)
    assert np.allclose(p["moving_variance"] + 1e-3, 0.75 + 1e-5, rtol=0, atol=1e-7)


def test_unsupported_architecture_fails_loudly(tmp_path):
    with pytest.raises(ValueError):
        convert.export(torch.nn.Linear(4, 4), str(tmp_path / "x.json"))
