"""Python mirror of the reference's public engine API over the C-ABI (numpy in, numpy out).

  GpuContext            ~ snn::GpuContext + dp::DeviceBackend            (core/inc/snn/snn.h, core/src/ic2/backend.h)
  ImageTexture          ~ snn::ImageTexture                              (core/inc/snn/imageTexture.h)
  MixedInferenceCore    ~ snn::MixedInferenceCore::create / run          (core/inc/snn/core.h:66-146)
  ShaderUnitTest-style single-layer helpers (conv2d, depthwise, ...)     (demo/common/shaderUnitTest.cpp:174-280)

Every call goes through libsnn_b200.so; nothing here computes.
"""
import ctypes as C
import sys

import numpy as np

from . import _lib
from ._lib import ConvDesc, ModelOptions, check, lib, vp

ACT = {"": 0, "linear": 0, "identity": 0, "none": 0, "relu": 1, "relu6": 2, "tanh": 3, "sigmoid": 4, "leakyRelu": 5, "leaky_relu": 5, "SiLU": 6,
       "softmax": 7}
PAD_MODE = {"": 0, "none": 0, "constant": 1, "replicate": 2, "reflect": 3}
ALGO = {"auto": 0, "simt": 1, "tcgen05": 2, "tcgen05-streamk": 3}
PRECISION = {"fp32x3": 0, "fp16w": 1, "fp16": 2}


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


class GpuContext:
    def __init__(self, device=0):
        self.h = vp()
        check(lib().snnb_context_create(int(device), C.byref(self.h)), "snnb_context_create")
        self.device = device

    def sync(self):
        check(lib().snnb_sync(self.h), "snnb_sync")

    def set_precision(self, precision):
        """Default product form of per-operator convolution launches: "fp32x3" (three bf16 MMAs) or "fp16w" (fp16 weights, two)."""
        check(lib().snnb_context_set_precision(self.h, PRECISION[precision]), "snnb_context_set_precision")

    @property
    def stream(self):
        return lib().snnb_context_stream(self.h)

    @property
    def launches(self):
        return int(lib().snnb_launch_count(self.h))

    def close(self):
        if self.h:
            lib().snnb_context_destroy(self.h)
            self.h = vp()

    def __del__(self):
        if sys.is_finalizing():  # interpreter shutdown: the CUDA runtime may already be gone, leave the handle to the OS
            return
        try:
            self.close()
        except Exception:
            pass


class ImageTexture:
    """Device tensor [N,H,W,C]; upload()/download() as in imageTexture.h, NHWC fp32 or the reference's C4HW4."""

    def __init__(self, ctx, n, h, w, c):
        self.ctx = ctx
        self.shape = (int(n), int(h), int(w), int(c))
        self.h = vp()
        check(lib().snnb_tensor_alloc(ctx.h, *self.shape, C.byref(self.h)), "snnb_tensor_alloc")

    @classmethod
    def from_numpy(cls, ctx, arr):
        arr = _f32(arr)
        assert arr.ndim == 4, "expect NHWC"
        t = cls(ctx, *arr.shape)
        t.upload(arr)
        return t

    def upload(self, arr):
        arr = _f32(arr)
        assert arr.shape == self.shape, (arr.shape, self.shape)
        check(lib().snnb_tensor_upload_nhwc(self.ctx.h, self.h, _ptr(arr)), "snnb_tensor_upload_nhwc")

    def download(self):
        out = np.empty(self.shape, dtype=np.float32)
        check(lib().snnb_tensor_download_nhwc(self.ctx.h, self.h, _ptr(out)), "snnb_tensor_download_nhwc")
        return out

    def upload_c4hw4(self, arr):
        n, h, w, c = self.shape
        arr = _f32(arr)
        assert arr.shape == (n, (c + 3) // 4, h, w, 4)
        check(lib().snnb_tensor_upload_c4hw4(self.ctx.h, self.h, _ptr(arr)), "snnb_tensor_upload_c4hw4")

    def download_c4hw4(self):
        n, h, w, c = self.shape
        out = np.empty((n, (c + 3) // 4, h, w, 4), dtype=np.float32)
        check(lib().snnb_tensor_download_c4hw4(self.ctx.h, self.h, _ptr(out)), "snnb_tensor_download_c4hw4")
        return out

    def dump(self, path):
        check(lib().snnb_tensor_dump(self.ctx.h, self.h, path.encode()), "snnb_tensor_dump")

    def free(self):
        if self.h:
            lib().snnb_tensor_free(self.h)
            self.h = vp()

    def __del__(self):
        if sys.is_finalizing():
            return
        try:
            self.free()
        except Exception:
            pass


class Weights:
    def __init__(self, handle):
        self.h = handle

    def free(self):
        if self.h:
            lib().snnb_weights_free(self.h)
            self.h = vp()

    def __del__(self):
        if sys.is_finalizing():
            return
        try:
            self.free()
        except Exception:
            pass


def conv_desc(ic, oc, k, stride=1, pad_x=0, pad_y=0, pad_mode="constant", activation="", alpha=0.0, algo="auto"):
    return ConvDesc(int(ic), int(oc), int(k), int(stride), int(pad_x), int(pad_y), PAD_MODE[pad_mode], ACT[activation], float(alpha), ALGO[algo])


def _bn_ptrs(bn):
    if bn is None:
        return [None] * 4, []
    keep = [_f32(bn[k]) for k in ("gamma", "beta", "mean", "var")]
    return [_ptr(a) for a in keep], keep


# ---- single-layer helpers (the boundary the reference's own op tests use: ShaderUnitTest::snn*TestWithLayer) ----
def conv2d(ctx, x, w_oihw, bias=None, bn=None, stride=1, pad_x=0, pad_y=0, pad_mode="constant", activation="", alpha=0.0, out_hw=None, residual=None,
           algo="auto"):
    x = _f32(x)
    w_oihw = _f32(w_oihw)
    n, h, w, ic = x.shape
    oc, ic2, k, _ = w_oihw.shape
    assert ic == ic2
    d = conv_desc(ic, oc, k, stride, pad_x, pad_y, pad_mode, activation, alpha, algo)
    bias_a = None if bias is None else _f32(bias)
    ptrs, keep = _bn_ptrs(bn)
    wh = vp()
    check(lib().snnb_weights_pack_conv2d(ctx.h, C.byref(d), _ptr(w_oihw), _ptr(bias_a), *ptrs, C.byref(wh)), "snnb_weights_pack_conv2d")
    wobj = Weights(wh)
    oh, ow = out_hw
    tin = ImageTexture.from_numpy(ctx, x)
    tout = ImageTexture(ctx, n, oh, ow, oc)
    tres = None if residual is None else ImageTexture.from_numpy(ctx, residual)
    check(lib().snnb_conv2d_launch(ctx.h, C.byref(d), wobj.h, tin.h, tres.h if tres else None, tout.h), "snnb_conv2d_launch")
    out = tout.download()
    del keep
    return out


def depthwise(ctx, x, w_chw, bias=None, bn=None, stride=1, pad_x=0, pad_y=0, activation="", alpha=0.0, out_hw=None):
    x = _f32(x)
    w_chw = _f32(w_chw)
    n, h, w, c = x.shape
    k = w_chw.shape[-1]
    d = conv_desc(c, c, k, stride, pad_x, pad_y, "constant", activation, alpha)
    bias_a = None if bias is None else _f32(bias)
    ptrs, keep = _bn_ptrs(bn)
    wh = vp()
    check(lib().snnb_weights_pack_depthwise(ctx.h, C.byref(d), _ptr(w_chw), _ptr(bias_a), *ptrs, C.byref(wh)), "snnb_weights_pack_depthwise")
    wobj = Weights(wh)
    oh, ow = out_hw
    tin = ImageTexture.from_numpy(ctx, x)
    tout = ImageTexture(ctx, n, oh, ow, c)
    check(lib().snnb_depthwise_launch(ctx.h, C.byref(d), wobj.h, tin.h, tout.h), "snnb_depthwise_launch")
    out = tout.download()
    del keep
    return out


def pool2d(ctx, x, k, stride, avg, out_hw):
    x = _f32(x)
    n, h, w, c = x.shape
    tin = ImageTexture.from_numpy(ctx, x)
    tout = ImageTexture(ctx, n, out_hw[0], out_hw[1], c)
    fn = lib().snnb_avgpool_launch if avg else lib().snnb_maxpool_launch
    check(fn(ctx.h, int(k), int(stride), tin.h, tout.h), "snnb_pool_launch")
    return tout.download()


def add(ctx, a, b, activation="", alpha=0.0):
    ta, tb = ImageTexture.from_numpy(ctx, a), ImageTexture.from_numpy(ctx, b)
    to = ImageTexture(ctx, *ta.shape)
    check(lib().snnb_add_launch(ctx.h, ACT[activation], float(alpha), ta.h, tb.h, to.h), "snnb_add_launch")
    return to.download()


def _channels(ctx, c, gamma=None, beta=None, mean=None, var=None):
    arrs = [None if a is None else _f32(a) for a in (gamma, beta, mean, var)]
    wh = vp()
    check(lib().snnb_weights_pack_channels(ctx.h, int(c), *[_ptr(a) for a in arrs], C.byref(wh)), "snnb_weights_pack_channels")
    return Weights(wh)


def batchnorm(ctx, x, bn, activation="", alpha=0.0):
    t = ImageTexture.from_numpy(ctx, x)
    w = _channels(ctx, t.shape[3], bn["gamma"], bn["beta"], bn["mean"], bn["var"])
    o = ImageTexture(ctx, *t.shape)
    check(lib().snnb_batchnorm_launch(ctx.h, w.h, ACT[activation], float(alpha), t.h, o.h), "snnb_batchnorm_launch")
    return o.download()


def instancenorm(ctx, x, gamma, beta, activation="", alpha=0.0):
    t = ImageTexture.from_numpy(ctx, x)
    w = _channels(ctx, t.shape[3], gamma, beta)
    o = ImageTexture(ctx, *t.shape)
    check(lib().snnb_instancenorm_launch(ctx.h, w.h, ACT[activation], float(alpha), t.h, o.h), "snnb_instancenorm_launch")
    return o.download()


def activation(ctx, x, activation, alpha=0.0):
    t = ImageTexture.from_numpy(ctx, x)
    o = ImageTexture(ctx, *t.shape)
    check(lib().snnb_activation_launch(ctx.h, ACT[activation], float(alpha), t.h, o.h), "snnb_activation_launch")
    return o.download()


def dense(ctx, x, kernel_out_in, bias=None, activation="", alpha=0.0):
    x = _f32(x)
    kernel_out_in = _f32(kernel_out_in)
    n_out, n_in = kernel_out_in.shape
    bias_a = None if bias is None else _f32(bias)
    wh = vp()
    check(lib().snnb_weights_pack_dense(ctx.h, n_in, n_out, _ptr(kernel_out_in), _ptr(bias_a), C.byref(wh)), "snnb_weights_pack_dense")
    w = Weights(wh)
    t = ImageTexture.from_numpy(ctx, x)
    o = ImageTexture(ctx, x.shape[0], 1, 1, n_out)
    check(lib().snnb_dense_launch(ctx.h, w.h, ACT[activation], float(alpha), t.h, o.h), "snnb_dense_launch")
    return o.download()


def softmax(ctx, x):
    t = ImageTexture.from_numpy(ctx, x)
    o = ImageTexture(ctx, *t.shape)
    check(lib().snnb_softmax_launch(ctx.h, t.h, o.h), "snnb_softmax_launch")
    return o.download()


def argmax1(ctx, x):
    t = ImageTexture.from_numpy(ctx, x)
    idx = (C.c_int * t.shape[0])()
    check(lib().snnb_argmax1(ctx.h, t.h, idx), "snnb_argmax1")
    return np.array(list(idx), dtype=np.int32)


def flatten(ctx, x):
    t = ImageTexture.from_numpy(ctx, x)
    n, h, w, c = t.shape
    o = ImageTexture(ctx, n, 1, 1, h * w * c)
    check(lib().snnb_flatten_launch(ctx.h, t.h, o.h), "snnb_flatten_launch")
    return o.download()


def concat(ctx, a, b):
    ta, tb = ImageTexture.from_numpy(ctx, a), ImageTexture.from_numpy(ctx, b)
    n, h, w, ca = ta.shape
    o = ImageTexture(ctx, n, h, w, ca + tb.shape[3])
    check(lib().snnb_concat_launch(ctx.h, ta.h, tb.h, o.h), "snnb_concat_launch")
    return o.download()


def upsample(ctx, x, scale, bilinear=False):
    t = ImageTexture.from_numpy(ctx, x)
    n, h, w, c = t.shape
    o = ImageTexture(ctx, n, int(h * scale), int(w * scale), c)
    check(lib().snnb_upsample_launch(ctx.h, float(scale), int(bool(bilinear)), t.h, o.h), "snnb_upsample_launch")
    return o.download()


def pad(ctx, x, pad_x, pad_y, out_hw, mode="constant"):
    t = ImageTexture.from_numpy(ctx, x)
    n, h, w, c = t.shape
    o = ImageTexture(ctx, n, out_hw[0], out_hw[1], c)
    check(lib().snnb_pad_launch(ctx.h, int(pad_x), int(pad_y), PAD_MODE[mode], t.h, o.h), "snnb_pad_launch")
    return o.download()


def subpixel(ctx, x, r):
    t = ImageTexture.from_numpy(ctx, x)
    n, h, w, c = t.shape
    o = ImageTexture(ctx, n, h * r, w * r, 1)
    check(lib().snnb_subpixel_launch(ctx.h, int(r), t.h, o.h), "snnb_subpixel_launch")
    return o.download()


class MixedInferenceCore:
    """snn::MixedInferenceCore: create(ctx, modelFileName, options) then run(images)."""

    def __init__(self, ctx, json_path, batch=1, input_hw=None, conv_algo="auto", use_cuda_graph=False, fuse=False, precision="fp32x3"):
        self.ctx = ctx
        self.batch = int(batch)
        opt = ModelOptions(self.batch, int(input_hw[1]) if input_hw else 0, int(input_hw[0]) if input_hw else 0, ALGO[conv_algo], int(bool(use_cuda_graph)),
                           int(bool(fuse)), PRECISION[precision])
        self.h = vp()
        check(lib().snnb_model_load_json(ctx.h, json_path.encode(), C.byref(opt), C.byref(self.h)), "snnb_model_load_json")

    @classmethod
    def create(cls, ctx, json_path, **kw):
        return cls(ctx, json_path, **kw)

    # -- introspection --
    @property
    def num_layers(self):
        return lib().snnb_model_num_layers(self.h)

    def layer_info(self, i):
        name = C.create_string_buffer(512)
        typ = C.create_string_buffer(64)
        d = [C.c_int() for _ in range(4)]
        check(lib().snnb_model_layer_info(self.h, i, name, 512, typ, 64, *[C.byref(v) for v in d]), "snnb_model_layer_info")
        return name.value.decode(), typ.value.decode(), tuple(v.value for v in d)

    def _dims(self, fn, idx):
        d = [C.c_int() for _ in range(4)]
        check(fn(self.h, idx, *[C.byref(v) for v in d]), "dims")
        return tuple(v.value for v in d)

    def input_shape(self, idx=0):
        return self._dims(lib().snnb_model_input_dims, idx)

    def output_shape(self, idx=0):
        return self._dims(lib().snnb_model_output_dims, idx)

    @property
    def num_outputs(self):
        return lib().snnb_model_num_outputs(self.h)

    @property
    def launches_per_forward(self):
        return lib().snnb_model_launches_per_forward(self.h)

    # -- execution --
    def run(self, images, want_classes=True):
        """End to end with host buffers: returns (output0 NHWC fp32, 1-based class indices or None)."""
        images = _f32(images)
        assert images.shape == self.input_shape(0), (images.shape, self.input_shape(0))
        oshape = self.output_shape(0)
        out = np.empty(oshape, dtype=np.float32)
        classes = (C.c_int * self.batch)()
        check(lib().snnb_model_run(self.h, _ptr(images), _ptr(out), out.size, classes if want_classes else None), "snnb_model_run")
        return out, (np.array(list(classes), dtype=np.int32) if want_classes else None)

    def run_raw(self, in_ptr, out_ptr, out_floats, classes_ptr=None):
        """snnb_model_run on raw host addresses (pinned buffers owned by the caller, e.g. bench.py)."""
        check(lib().snnb_model_run(self.h, in_ptr, out_ptr, out_floats, classes_ptr), "snnb_model_run")

    def submit_raw(self, in_ptr, out_ptr, out_floats, classes_ptr=None):
        """snnb_model_submit on raw (pinned) host addresses; returns the ticket for wait()."""
        t = C.c_int()
        check(lib().snnb_model_submit(self.h, in_ptr, out_ptr, out_floats, classes_ptr, C.byref(t)), "snnb_model_submit")
        return t.value

    def submit_u8_raw(self, in_ptr, mean4, norm4, out_ptr, out_floats, classes_ptr=None):
        """snnb_model_submit_u8: 8-bit NHWC images, normalised on the device as (x - mean[c & 3]) * norm[c & 3]."""
        t = C.c_int()
        m = (C.c_float * 4)(*[float(v) for v in mean4])
        n = (C.c_float * 4)(*[float(v) for v in norm4])
        check(lib().snnb_model_submit_u8(self.h, in_ptr, m, n, out_ptr, out_floats, classes_ptr, C.byref(t)), "snnb_model_submit_u8")
        return t.value

    def run_u8(self, images_u8, mean4, norm4, want_classes=True):
        """Convenience wrapper: one batch of uint8 NHWC images through submit_u8 / wait (numpy in, numpy out)."""
        images_u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
        assert images_u8.shape == self.input_shape(0), (images_u8.shape, self.input_shape(0))
        out = np.empty(self.output_shape(0), np.float32)
        classes = (C.c_int * self.batch)() if want_classes else None
        t = self.submit_u8_raw(images_u8.ctypes.data, mean4, norm4, out.ctypes.data, out.size, classes)
        self.wait(t)
        return out, (np.array(list(classes), dtype=np.int32) if want_classes else None)

    def run_image(self, images_u8, mean4, norm4, linear=True, out_u8=False, out_scale=1.0, out_offset=0.0, want_classes=False):
        """snnb_model_submit_image + wait: u8 NHWC images of ANY size in (resized + normalised on the device), fp32 tensor or u8 image out."""
        from ._lib import ImageIO
        images_u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
        n, sh, sw, c = images_u8.shape
        oshape = self.output_shape(0)
        out = np.empty(oshape, np.uint8 if out_u8 else np.float32)
        classes = (C.c_int * self.batch)() if want_classes else None
        io = ImageIO()
        io.input_u8, io.src_height, io.src_width, io.linear_filter = images_u8.ctypes.data, sh, sw, int(bool(linear))
        io.mean4 = (C.c_float * 4)(*[float(v) for v in mean4])
        io.norm4 = (C.c_float * 4)(*[float(v) for v in norm4])
        io.output_capacity = out.size
        if out_u8:
            io.output_u8, io.out_scale, io.out_offset = out.ctypes.data, float(out_scale), float(out_offset)
        else:
            io.output_f32 = out.ctypes.data
        io.classes_1based = C.cast(classes, C.c_void_p) if want_classes else None
        t = C.c_int()
        check(lib().snnb_model_submit_image(self.h, C.byref(io), C.byref(t)), "snnb_model_submit_image")
        self.wait(t.value)
        return out, (np.array(list(classes), dtype=np.int32) if want_classes else None)

    def wait(self, ticket):
        check(lib().snnb_model_wait(self.h, int(ticket)), "snnb_model_wait")

    def set_input(self, images, idx=0):
        images = _f32(images)
        check(lib().snnb_model_set_input(self.h, idx, _ptr(images)), "snnb_model_set_input")
        self.ctx.sync()

    def forward(self):
        check(lib().snnb_model_forward(self.h), "snnb_model_forward")

    def get_output(self, idx=0):
        out = np.empty(self.output_shape(idx), dtype=np.float32)
        check(lib().snnb_model_get_output(self.h, idx, _ptr(out), out.size), "snnb_model_get_output")
        return out

    def layer_output(self, layer):
        _, _, shape = self.layer_info(layer)
        out = np.empty(shape, dtype=np.float32)
        check(lib().snnb_model_layer_output(self.h, layer, _ptr(out), out.size), "snnb_model_layer_output")
        return out

    def time_layers(self):
        n = self.num_layers
        arr = (C.c_float * n)()
        check(lib().snnb_model_time_layers(self.h, arr, n), "snnb_model_time_layers")
        return np.array(list(arr), dtype=np.float32)

    def layer_kernel(self, layer):
        """Kernel launched by `layer` in the last time_layers() pass ("" if none)."""
        buf = C.create_string_buffer(128)
        check(lib().snnb_model_layer_kernel(self.h, int(layer), buf, 128), "snnb_model_layer_kernel")
        return buf.value.decode()

    def dump_outputs(self, directory):
        check(lib().snnb_model_dump_outputs(self.h, directory.encode()), "snnb_model_dump_outputs")

    def boxes(self, n=0, max_rows=100):
        rows = (C.c_float * (6 * max_rows))()
        cnt = C.c_int()
        check(lib().snnb_model_get_boxes(self.h, n, rows, max_rows, C.byref(cnt)), "snnb_model_get_boxes")
        return np.array(list(rows), dtype=np.float32).reshape(max_rows, 6)[:cnt.value]

    def weight_arena(self):
        p = vp()
        b = C.c_size_t()
        check(lib().snnb_model_weight_arena(self.h, C.byref(p), C.byref(b)), "snnb_model_weight_arena")
        return p.value, b.value

    def close(self):
        if self.h:
            lib().snnb_model_destroy(self.h)
            self.h = vp()

    def __del__(self):
        if sys.is_finalizing():  # interpreter shutdown: the CUDA runtime may already be gone, leave the handle to the OS
            return
        try:
            self.close()
        except Exception:
            pass
