"""torch / torchvision -> ShaderNN JSON model writer (SURVEY §8 f-N1).

The reference ships `tools/convertTool` (Keras h5 / ONNX -> JSON, `convertProcessor/converters/h5ToJsonConverter.py:74-190`,
`onnxToJsonConverter.py:34-278`); its modelzoo JSONs are LFS stubs in the checkout, so "models drop in unchanged" needs a way
to produce real files. This module writes the same format (through `modelzoo.write_model`, embedded or `_layers.json` +
`_weights.bin`) from live `torch.nn` modules, applying the converter's reorder rules and the engine's quirks:

* Conv2d weight `[O][I][kh][kw]` is already OIHW (`modelparser.cpp:641-657`); Linear weight `[out][in]` is already the Dense
  layout (`cpulayer.h:162`); depthwise weight `[C][1][kh][kw]` -> canonical `[C][kh][kw]` (the writer emits the HWC-major flat
  list for JSON and CHW for `.bin`, `modelparser.cpp:827-850`).
* BatchNorm eps: the reference hard-codes `sqrt(var + 1e-3)` (`shadertemplate_vk_conv2d.comp:282-283`); a torch module with
  another eps is represented exactly by writing `moving_variance = var + eps - 1e-3`.
* torch's padded max pool (`MaxPool2d(3, 2, padding=1)`): reference pools never pad top/left
  (`maxpool2dVulkan.cpp:57-60`), so an explicit `ZeroPadding2D ((1,0),(1,0))` + `valid` pool is emitted; zero padding is exact
  behind a ReLU (inputs >= 0).
* symmetric conv padding p = k//2 is `"same"`, p = 0 is `"valid"`; anything else is written numerically `[[p,p],[p,p]]`.

Only module types on the hot path are handled (ResNet BasicBlock nets, MobileNetV2); the walkers are explicit about the
architecture instead of tracing, so an unsupported module fails loudly.
"""
import numpy as np

from . import modelzoo


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


class TorchExporter(modelzoo.Builder):
    """modelzoo.Builder whose layers take their weights from torch modules instead of the synthetic distributions."""

    def _padding(self, conv):
        k, p = conv.kernel_size[0], conv.padding[0]
        assert conv.kernel_size[0] == conv.kernel_size[1] and conv.padding[0] == conv.padding[1] and conv.stride[0] == conv.stride[1], "square kernels only"
        assert conv.dilation[0] == 1, "dilation is fixed at 1 in the reference (conv2dVulkan.cpp:89)"
        if p == 0:
            return "valid"
        if p == k // 2 and k % 2 == 1:
            return "same"
        return [[p, p], [p, p]]

    @staticmethod
    def _bn_params(bn):
        return {"gamma": _np(bn.weight), "beta": _np(bn.bias), "moving_mean": _np(bn.running_mean),
                "moving_variance": _np(bn.running_var) + np.float32(bn.eps) - np.float32(1e-3)}  # reference eps is fixed at 1e-3

    def conv_from(self, x, conv, bn=None, activation="linear"):
        assert conv.groups == 1
        i = self.conv(x, conv.out_channels, conv.kernel_size[0], conv.stride[0], self._padding(conv), activation, bias=conv.bias is not None, bn=bn is not None)
        d = self.layers[i]
        assert d["inputPlanes"] == conv.in_channels, (d["inputPlanes"], conv.in_channels)
        d["_w"] = {"kernel": _np(conv.weight)}
        if conv.bias is not None:
            d["_w"]["bias"] = _np(conv.bias)
        if bn is not None:
            d["_bn"] = self._bn_params(bn)
        return i

    def depthwise_from(self, x, conv, bn=None, activation="linear"):
        assert conv.groups == conv.in_channels == conv.out_channels, "depth multiplier 1 only (modelparser.cpp:821)"
        i = self.depthwise(x, conv.kernel_size[0], conv.stride[0], self._padding(conv), activation, bias=conv.bias is not None, bn=bn is not None)
        d = self.layers[i]
        d["_w"] = {"kernel_chw": _np(conv.weight)[:, 0]}
        if conv.bias is not None:
            d["_w"]["bias"] = _np(conv.bias)
        if bn is not None:
            d["_bn"] = self._bn_params(bn)
        return i

    def dense_from(self, x, linear, activation="linear"):
        i = self.dense(x, linear.in_features, linear.out_features, activation, bias=linear.bias is not None)
        d = self.layers[i]
        d["_w"] = {"kernel": _np(linear.weight)}
        if linear.bias is not None:
            d["_w"]["bias"] = _np(linear.bias)
        return i

    def maxpool_from(self, x, pool):
        k = pool.kernel_size if isinstance(pool.kernel_size, int) else pool.kernel_size[0]
        s = pool.stride if isinstance(pool.stride, int) else pool.stride[0]
        p = pool.padding if isinstance(pool.padding, int) else pool.padding[0]
        assert not pool.ceil_mode and p <= k // 2
        if p:
            x = self.pad(x, p, 0, p, 0)  # bottom/right overhang is clipped by the pool itself
        return self.maxpool(x, k, s, "valid")


def from_torchvision_resnet(model, input_hw=(224, 224)):
    """torchvision.models.resnet18 / resnet34 (BasicBlock) in eval mode -> layer list."""
    b = TorchExporter()
    h, w = input_hw
    x = b.input(w, h, 3)
    x = b.conv_from(x, model.conv1, model.bn1, "relu")
    x = b.maxpool_from(x, model.maxpool)
    for stage in (model.layer1, model.layer2, model.layer3, model.layer4):
        for blk in stage:
            assert type(blk).__name__ == "BasicBlock", "only BasicBlock ResNets are on the hot path"
            y = b.conv_from(x, blk.conv1, blk.bn1, "relu")
            y = b.conv_from(y, blk.conv2, blk.bn2, "linear")
            s = x if blk.downsample is None else b.conv_from(x, blk.downsample[0], blk.downsample[1], "linear")
            x = b.add(y, s, "relu")
    feat = b.planes(x)
    assert h == w and h % 32 == 0, "the global pool is written as a square AveragePooling2D(pool=[H,H]) (averagepooling2d.py:40-55)"
    x = b.global_avgpool(x, max(1, h // 32))
    x = b.flatten(x, feat)
    x = b.dense_from(x, model.fc, "linear")
    return b.layers


def from_torchvision_mobilenet_v2(model, input_hw=(224, 224)):
    """torchvision.models.mobilenet_v2 in eval mode -> layer list."""
    b = TorchExporter()
    h, w = input_hw
    x = b.input(w, h, 3)
    feats = list(model.features)
    x = b.conv_from(x, feats[0][0], feats[0][1], "relu6")
    for blk in feats[1:-1]:
        assert type(blk).__name__ == "InvertedResidual"
        mods = list(blk.conv)
        y = x
        if len(mods) == 4:  # expand 1x1
            y = b.conv_from(y, mods[0][0], mods[0][1], "relu6")
            mods = mods[1:]
        y = b.depthwise_from(y, mods[0][0], mods[0][1], "relu6")
        y = b.conv_from(y, mods[1], mods[2], "linear")
        x = b.add(y, x, "linear") if blk.use_res_connect else y
    x = b.conv_from(x, feats[-1][0], feats[-1][1], "relu6")
    feat = b.planes(x)
    assert h == w and h % 32 == 0, "the global pool is written as a square AveragePooling2D(pool=[H,H]) (averagepooling2d.py:40-55)"
    x = b.global_avgpool(x, max(1, h // 32))
    x = b.flatten(x, feat)
    x = b.dense_from(x, model.classifier[-1], "linear")
    return b.layers


def export(model, path, input_hw=(224, 224), split=False):
    """Write `model` (a supported torchvision architecture, eval mode) as a ShaderNN JSON model; returns the layer list."""
    name = type(model).__name__
    if name == "ResNet":
        layers = from_torchvision_resnet(model, input_hw)
    elif name == "MobileNetV2":
        layers = from_torchvision_mobilenet_v2(model, input_hw)
    else:
        raise ValueError("convert.export: unsupported architecture %s (supported: torchvision ResNet[BasicBlock], MobileNetV2)" % name)
    modelzoo.write_model(layers, path, split=split)
    return layers
