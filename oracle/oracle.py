"""TEST INFRASTRUCTURE — Python face of the parity oracle (oracle/libsnn_oracle.so) plus a whole-model walker.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this. It reads
the SNN JSON (+ sidecar .bin) model with Python's own json module — an implementation independent of the product's
C++ ModelParser — and evaluates the graph layer by layer with the oracle's C++ operators (NHWC fp32 numpy arrays).
"""
import ctypes as C
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsnn_oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libsnn_ref.so")

_lib = None
_ref = None
fp = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s missing: run `make -C oracle` (or __graft_entry__.build())" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        l.orc_rand_u64.restype = C.c_uint64
        l.orc_random_float.restype = C.c_float
        l.orc_random_float.argtypes = [C.c_float, C.c_float]
        l.orc_srand.argtypes = [C.c_uint64]
        l.orc_random_fill.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float]
        l.orc_compare.restype = C.c_size_t
        l.orc_compare.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float]
        l.orc_add.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_void_p]
        l.orc_activation.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_void_p]
        l.orc_batchnorm.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        l.orc_flatten.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        l.orc_concat.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
        l.orc_conv2d.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int] * 7 + [C.c_float, C.c_void_p, C.c_int, C.c_int]
        l.orc_depthwise.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_float, C.c_void_p, C.c_int, C.c_int]
        l.orc_pool2d.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_int, C.c_int]
        l.orc_dense.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        l.orc_softmax.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        l.orc_argmax1.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        l.orc_upsample.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_int]
        l.orc_pad.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p, C.c_int, C.c_int]
        l.orc_instancenorm.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        l.orc_subpixel.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]
        l.orc_yolo.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_int]
        l.orc_hwc_to_c4hw4.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.orc_c4hw4_to_hwc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = l
    return _lib


def ref():
    """The compiled slice of the reference itself (oracle/_ref): cpulayer.h Dense/activations + prng.h."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_PATH):
            return None
        r = C.CDLL(REF_PATH)
        r.ref_rand_u64.restype = C.c_uint64
        r.ref_srand.argtypes = [C.c_uint64]
        r.ref_random_float.restype = C.c_float
        r.ref_random_float.argtypes = [C.c_float, C.c_float]
        r.ref_dense.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_float, C.c_void_p]
        r.ref_activation.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_float, C.c_void_p]
        _ref = r
    return _ref


ACT = {"": 0, "linear": 0, "identity": 0, "none": 0, "relu": 1, "relu6": 2, "tanh": 3, "sigmoid": 4, "leakyRelu": 5, "leaky_relu": 5, "SiLU": 6,
       "softmax": 7}
PAD_MODE = {"": 0, "none": 0, "constant": 1, "replicate": 2, "reflect": 3}


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _bn_pack(bn):
    """dict(gamma,beta,mean,var) -> one [4*C] array in the oracle's order."""
    if bn is None:
        return None
    return _f(np.concatenate([_f(bn["gamma"]), _f(bn["beta"]), _f(bn["mean"]), _f(bn["var"])]))


# ---- dims ----
def conv_out_dim(n, k, s, pad_t, pad_b):
    return lib().orc_conv_out_dim(int(n), int(k), int(s), int(pad_t), int(pad_b))


def depthwise_out_dim(n, k, s, pa, pb):
    return lib().orc_depthwise_out_dim(int(n), int(k), int(s), int(pa), int(pb))


def pool_out_dim(n, k, s, valid_like):
    return lib().orc_pool_out_dim(int(n), int(k), int(s), int(bool(valid_like)))


def same_padding(k, is_same):
    offs = (C.c_int * 4)()
    lib().orc_same_padding(int(k), int(bool(is_same)), offs)
    return list(offs)


# ---- operators (NHWC fp32 numpy) ----
def conv2d(x, w_oihw, bias=None, bn=None, stride=1, pad_x=0, pad_y=0, pad_mode="constant", activation="", alpha=0.0, out_hw=None):
    x, w = _f(x), _f(w_oihw)
    n, h, wd, ic = x.shape
    oc, _, k, _ = w.shape
    oh, ow = out_hw
    y = np.empty((n, oh, ow, oc), np.float32)
    b = None if bias is None else _f(bias)
    bnp = _bn_pack(bn)
    lib().orc_conv2d(_p(x), n, h, wd, ic, _p(w), _p(b), _p(bnp), oc, k, int(stride), int(pad_x), int(pad_y), PAD_MODE[pad_mode], ACT[activation],
                     float(alpha), _p(y), oh, ow)
    return y


def depthwise(x, w_chw, bias=None, bn=None, stride=1, pad_x=0, pad_y=0, activation="", alpha=0.0, out_hw=None):
    x, w = _f(x), _f(w_chw)
    n, h, wd, c = x.shape
    k = w.shape[-1]
    oh, ow = out_hw
    y = np.empty((n, oh, ow, c), np.float32)
    b = None if bias is None else _f(bias)
    bnp = _bn_pack(bn)
    lib().orc_depthwise(_p(x), n, h, wd, c, _p(w), _p(b), _p(bnp), k, int(stride), int(pad_x), int(pad_y), ACT[activation], float(alpha), _p(y), oh, ow)
    return y


def pool2d(x, k, stride, avg, out_hw):
    x = _f(x)
    n, h, w, c = x.shape
    y = np.empty((n, out_hw[0], out_hw[1], c), np.float32)
    lib().orc_pool2d(_p(x), n, h, w, c, int(k), int(stride), int(bool(avg)), _p(y), out_hw[0], out_hw[1])
    return y


def add(a, b, activation="", alpha=0.0):
    a, b = _f(a), _f(b)
    assert a.shape == b.shape, ("Add: operand shapes differ", a.shape, b.shape)
    y = np.empty_like(a)
    lib().orc_add(_p(a), _p(b), a.size, ACT[activation], float(alpha), _p(y))
    return y


def batchnorm(x, bn, activation="", alpha=0.0):
    x = _f(x)
    y = np.empty_like(x)
    bnp = _bn_pack(bn)
    lib().orc_batchnorm(_p(x), x.size // x.shape[-1], x.shape[-1], _p(bnp), ACT[activation], float(alpha), _p(y))
    return y


def activation(x, act, alpha=0.0):
    x = _f(x)
    y = np.empty_like(x)
    lib().orc_activation(_p(x), x.size, ACT[act], float(alpha), _p(y))
    return y


def dense(x, kernel_out_in, bias=None, activation="", alpha=0.0):
    x, k = _f(x), _f(kernel_out_in)
    n = x.shape[0]
    n_out, n_in = k.shape
    x2 = x.reshape(n, n_in)
    y = np.empty((n, n_out), np.float32)
    b = None if bias is None else _f(bias)
    lib().orc_dense(_p(x2), n, n_in, _p(k), _p(b), n_out, ACT[activation], float(alpha), _p(y))
    return y.reshape(n, 1, 1, n_out)


def softmax(x):
    x = _f(x)
    y = np.empty_like(x)
    lib().orc_softmax(_p(x), x.size // x.shape[-1], x.shape[-1], _p(y))
    return y


def argmax1(x):
    x = _f(x)
    n = x.shape[0]
    idx = (C.c_int * n)()
    lib().orc_argmax1(_p(x.reshape(n, -1)), n, x.size // n, idx)
    return np.array(list(idx), np.int32)


def flatten(x):
    x = _f(x)
    return x.reshape(x.shape[0], 1, 1, -1).copy()


def concat(a, b):
    a, b = _f(a), _f(b)
    y = np.empty(a.shape[:3] + (a.shape[3] + b.shape[3],), np.float32)
    lib().orc_concat(_p(a), a.shape[3], _p(b), b.shape[3], a.size // a.shape[3], _p(y))
    return y


def upsample(x, scale, bilinear=False):
    x = _f(x)
    n, h, w, c = x.shape
    oh, ow = int(h * scale), int(w * scale)
    y = np.empty((n, oh, ow, c), np.float32)
    lib().orc_upsample(_p(x), n, h, w, c, float(scale), int(bool(bilinear)), _p(y), oh, ow)
    return y


def pad(x, pad_x, pad_y, out_hw, mode="constant"):
    x = _f(x)
    n, h, w, c = x.shape
    y = np.empty((n, out_hw[0], out_hw[1], c), np.float32)
    lib().orc_pad(_p(x), n, h, w, c, int(pad_x), int(pad_y), PAD_MODE[mode], _p(y), out_hw[0], out_hw[1])
    return y


def instancenorm(x, gamma, beta, activation="", alpha=0.0):
    x = _f(x)
    n, h, w, c = x.shape
    y = np.empty_like(x)
    g, b = _f(gamma), _f(beta)
    lib().orc_instancenorm(_p(x), n, h, w, c, _p(g), _p(b), ACT[activation], float(alpha), _p(y))
    return y


def subpixel(x, r):
    x = _f(x)
    n, h, w, c = x.shape
    y = np.empty((n, h * r, w * r, 1), np.float32)
    lib().orc_subpixel(_p(x), n, h, w, int(r), _p(y))
    return y


def yolo(head0, head1, net_hw=(416, 416), conf=0.35, iou=0.45):
    """head arrays [H,W,C] of ONE image -> rows {class, score, x, y, w, h}."""
    h0, h1 = _f(head0), _f(head1)
    out = np.zeros((100, 6), np.float32)
    n = lib().orc_yolo(_p(h0), _p(h1), h0.shape[-1], int(net_hw[1]), int(net_hw[0]), float(conf), float(iou), _p(out), 100)
    return out[:n]


def resize_normalize(images_u8, out_hw, mean4, norm4, linear=True):
    """ImageTexture::resize + normalisation (core/inc/snn/imageTexture.h:137; shadertemplate_vk_resize.comp:42-61): output texel
    (x, y) samples the source at ((x + 0.5) / outW, (y + 0.5) / outH) with the sampler's filter - linear: texel centres at
    (i + 0.5) / inW, weights from the fractional part, clamp to edge; nearest: floor - then (v - mean[c & 3]) * norm[c & 3].
    fp32 arithmetic in the order of the CUDA kernel it checks (bilinear as two lerps along x, then one along y)."""
    img = np.asarray(images_u8, np.uint8).astype(np.float32)
    n, sh, sw, c = img.shape
    oh, ow = out_hw
    f32 = np.float32
    fx = (np.arange(ow, dtype=f32) + f32(0.5)) / f32(ow) * f32(sw)
    fy = (np.arange(oh, dtype=f32) + f32(0.5)) / f32(oh) * f32(sh)
    if linear:
        sx, sy = fx - f32(0.5), fy - f32(0.5)
        x0f, y0f = np.floor(sx), np.floor(sy)
        ax, ay = (sx - x0f).astype(f32), (sy - y0f).astype(f32)
        x0, x1 = np.clip(x0f.astype(int), 0, sw - 1), np.clip(x0f.astype(int) + 1, 0, sw - 1)
        y0, y1 = np.clip(y0f.astype(int), 0, sh - 1), np.clip(y0f.astype(int) + 1, 0, sh - 1)
        ax_, ay_ = ax[None, None, :, None], ay[None, :, None, None]
        p00, p01 = img[:, y0][:, :, x0], img[:, y0][:, :, x1]
        p10, p11 = img[:, y1][:, :, x0], img[:, y1][:, :, x1]
        top = p00 + ax_ * (p01 - p00)
        bot = p10 + ax_ * (p11 - p10)
        v = top + ay_ * (bot - top)
    else:
        x0 = np.minimum(fx.astype(int), sw - 1)
        y0 = np.minimum(fy.astype(int), sh - 1)
        v = img[:, y0][:, :, x0]
    mean = np.array([mean4[i & 3] for i in range(c)], f32)
    norm = np.array([norm4[i & 3] for i in range(c)], f32)
    return ((v.astype(f32) - mean) * norm).astype(f32)


def compare(a, b, eps):
    """Number of mismatches under the reference's comparator (demo/common/testutil.cpp:351-361)."""
    a, b = _f(a).ravel(), _f(b).ravel()
    assert a.size == b.size
    return int(lib().orc_compare(_p(a), _p(b), a.size, float(eps)))


def random_mat(shape, lo=-1.2, hi=1.2):
    """RandomMat of the reference's tests (testutil.cpp:42-47) from the shared PRNG stream."""
    a = np.empty(shape, np.float32)
    lib().orc_random_fill(_p(a), a.size, float(lo), float(hi))
    return a


def srand(seed=7767517):
    lib().orc_srand(C.c_uint64(seed))


# ----------------------------------------------------------------------------------------------------------------
# model reader + walker (independent of the product's C++ parser)
# ----------------------------------------------------------------------------------------------------------------
class Model:
    def __init__(self, json_path):
        with open(json_path) as f:
            root = json.load(f)
        self.count = int(root["numLayers"]["count"])
        self.layers = [root["Layer_%d" % i] for i in range(self.count)]
        self.bin = None
        bname = root["numLayers"].get("bin_file_name")
        if bname:
            self.bin = np.fromfile(os.path.join(os.path.dirname(os.path.abspath(json_path)), bname), dtype="<f4")
            self.bin_pos = 0
        self._load_weights()

    def _take(self, n):
        a = self.bin[self.bin_pos:self.bin_pos + n]
        assert a.size == n, "sidecar .bin truncated"
        self.bin_pos += n
        return a.astype(np.float32)

    @staticmethod
    def type_of(l):
        t = l["type"]
        if t == "Lambda":
            t = l["name"]
        return {"DepthwiseConv2D": "SeparableConv2D", "Depthwise": "SeparableConv2D", "InstanceNormalization": "InstanceNorm", "ZeroPadding2D": "Pad",
                "subpixel": "Subpixel", "depth_to_space": "Subpixel"}.get(t, t)

    def _bn(self, l, c, allow_bin=True):
        if self.bin is not None and allow_bin:
            g, b, m, v = (self._take(c) for _ in range(4))
        else:
            o = l["batchNormalization"]
            g = np.array(o.get("gamma", [1.0] * c), np.float32)
            b = np.array(o.get("beta", [0.0] * c), np.float32)
            m = np.array(o["moving_mean"] if "moving_mean" in o else o["movingMean"], np.float32)
            v = np.array(o["moving_variance"] if "moving_variance" in o else o["movingVariance"], np.float32)
        return {"gamma": g, "beta": b, "mean": m, "var": v}

    def _load_weights(self):
        self.w = [None] * self.count
        for i, l in enumerate(self.layers):
            t = self.type_of(l)
            d = {}
            if t == "Conv2D":
                oc, ic, k = int(l["outputPlanes"]), int(l["inputPlanes"]), int(l["kernel_size"])
                n = oc * ic * k * k
                d["kernel"] = (self._take(n) if self.bin is not None else np.array(l["weights"]["kernel"], np.float32)).reshape(oc, ic, k, k)
                if l.get("useBias") == "True":
                    d["bias"] = self._take(oc) if self.bin is not None else np.array(l["weights"]["bias"], np.float32)
                if l.get("useBatchNormalization") == "True":
                    d["bn"] = self._bn(l, oc)
            elif t == "SeparableConv2D":
                c, k = int(l["inputPlanes"]), int(l["kernel_size"])
                if self.bin is not None:
                    d["kernel"] = self._take(c * k * k).reshape(c, k, k)
                else:
                    d["kernel"] = np.array(l["weights"]["kernel"], np.float32).reshape(k * k, c).T.reshape(c, k, k).copy()
                if l.get("useBias") == "True":
                    d["bias"] = self._take(c) if self.bin is not None else np.array(l["weights"]["bias"], np.float32)
                if l.get("useBatchNormalization") == "True":
                    d["bn"] = self._bn(l, c)
            elif t == "Dense":
                units = int(l.get("units", l["outputPlanes"]))
                if self.bin is not None:
                    n_in = int(l["inputPlanes"])
                    d["kernel"] = self._take(n_in * units).reshape(units, n_in)
                else:
                    flat = np.array(l["weights"]["kernel"], np.float32)
                    d["kernel"] = flat.reshape(units, flat.size // units)
                if l.get("useBias") == "True":
                    d["bias"] = self._take(units) if self.bin is not None else np.array(l["weights"]["bias"], np.float32)
            elif t == "BatchNormalization":
                d["bn"] = self._bn(l, int(l["outputPlanes"]), allow_bin=False)
            elif t == "InstanceNorm":
                d["gamma"] = np.array(l["weights"]["scale"], np.float32)
                d["beta"] = np.array(l["weights"]["bias"], np.float32)
            self.w[i] = d

    # padding spec -> [T, B, L, R]
    @staticmethod
    def _offsets(l, k, even_minus_one=True, key="padding"):
        p = l.get(key)
        if isinstance(p, list):
            if isinstance(p[0], list):
                return [int(p[0][0]), int(p[0][1]), int(p[1][0]), int(p[1][1])]
            return [int(p[0]), int(p[0]), int(p[1]), int(p[1])]
        if isinstance(p, (int, float)):
            return [int(p)] * 4
        if p in ("valid", "none"):
            return [0, 0, 0, 0]
        if k > 1:
            o = [max(k // 2, 1)] * 4
            if even_minus_one and k % 2 == 0:
                o[0] -= 1
                o[2] -= 1
            return o
        return [0, 0, 0, 0]

    def run(self, x, return_all=False, input_index=0):
        """Evaluate the graph on NHWC fp32 input x. Returns the last layer's output (or every layer's)."""
        outs = [None] * self.count
        done = [False] * self.count
        pending = list(range(self.count))
        boxes = None
        while pending:
            progressed = False
            for i in list(pending):
                l = self.layers[i]
                ins = [int(v) for v in l.get("inputId", [])][:int(l.get("numInputs", 0))]
                if any(not done[j] for j in ins):
                    continue
                outs[i] = self._eval(i, l, [outs[j] for j in ins], x)
                if isinstance(outs[i], tuple):
                    boxes = outs[i][1]
                    outs[i] = outs[i][0]
                done[i] = True
                pending.remove(i)
                progressed = True
            assert progressed, "cycle in model graph"
        self.boxes = boxes
        return outs if return_all else outs[-1]

    def _eval(self, i, l, ins, x):
        t = self.type_of(l)
        w = self.w[i]
        act = l.get("activation", "linear")
        if t == "InputLayer":
            return _f(x)
        a = ins[0]
        n, h, wd, c = a.shape
        if t == "Conv2D":
            k, s = int(l["kernel_size"]), int(l["strides"])
            o = self._offsets(l, k)
            oh, ow = conv_out_dim(h, k, s, o[0], o[1]), conv_out_dim(wd, k, s, o[0], o[1])
            alpha = float(l.get("leakyReluAlpha", l.get("alpha", 0.0)))
            mode = l.get("mode", "") if isinstance(l.get("padding"), list) and isinstance(l["padding"][0], list) else ""
            px, py = (0, 0) if k == 1 else (o[0], o[2])  # uPadx <- T, uPady <- L (conv2dVulkan.cpp:183-184)
            return conv2d(a, w["kernel"], w.get("bias"), w.get("bn"), s, px, py, mode, act, alpha, (oh, ow))
        if t == "SeparableConv2D":
            k, s = int(l["kernel_size"]), int(l["strides"])
            o = self._offsets(l, k)
            ow = depthwise_out_dim(wd, k, s, o[0], o[2])  # separableconvolution.cpp:77-86: width uses T+L
            oh = depthwise_out_dim(h, k, s, o[1], o[3])   #                                height uses B+R
            alpha = float(l.get("leakyReluAlpha", l.get("alpha", 0.0)))
            return depthwise(a, w["kernel"], w.get("bias"), w.get("bn"), s, o[0], o[2], act, alpha, (oh, ow))
        if t in ("MaxPooling2D", "AveragePooling2D"):
            pool = l.get("pool", l.get("pool_size"))
            k = int(pool[0] if isinstance(pool, list) else pool)
            # max pool reads "stride" or "strides"; avg pool ONLY "stride" (modelparser.cpp:312-335 vs :385-391)
            s = l.get("stride", l.get("strides", k)) if t == "MaxPooling2D" else l.get("stride", k)
            s = int(s[0] if isinstance(s, list) else s)
            pd = l.get("padding")
            pstr = str(int(pd)) if isinstance(pd, (int, float)) else (pd if isinstance(pd, str) else str(int(pd[0][0] if isinstance(pd[0], list) else pd[0])))
            valid_like = pstr in ("0", "valid", "none")
            oh, ow = pool_out_dim(h, k, s, valid_like), pool_out_dim(wd, k, s, valid_like)
            return pool2d(a, k, s, t == "AveragePooling2D", (oh, ow))
        if t == "Add":
            alpha = float(l.get("leakyReluAlpha", l.get("alpha", 0.3)))
            return add(a, ins[1], act, alpha)
        if t == "BatchNormalization":
            return batchnorm(a, w["bn"], act, float(l.get("leakyReluAlpha", l.get("alpha", 0.0))))
        if t == "InstanceNorm":
            return instancenorm(a, w["gamma"], w["beta"], act, float(l.get("leakyReluAlpha", 0.0)))
        if t == "Activation":
            return activation(a, act, float(l.get("leakyReluAlpha", 0.3)))
        if t == "Flatten":
            y = flatten(a)
            if act == "softmax":
                y = softmax(y)
            elif ACT.get(act, 0):
                y = activation(y, act)
            return y
        if t == "Dense":
            # NB the CPU reference's SiLU is a no-op (cpulayer.h:245-252); the product applies the real one
            y = dense(a, w["kernel"], w.get("bias"), "" if act == "SiLU" else act, float(l.get("leakyReluAlpha", l.get("alpha", 0.3))))
            return activation(y, "SiLU") if act == "SiLU" else y
        if t == "Concatenate":
            return concat(a, ins[1])
        if t == "UpSampling2D":
            return upsample(a, float(l["scaleFactor"]), l.get("interpolation") == "bilinear")
        if t == "Pad":
            if "pads" in l:
                p = l["pads"]
                o = [int(p[2]), int(p[6]), int(p[3]), int(p[7])]
            else:
                o = self._offsets(l, 0, False)
            oh, ow = h + o[0] + o[1], wd + o[2] + o[3]
            return pad(a, o[0], o[2], (oh, ow), l.get("mode", "constant"))  # uPad = (T, L)
        if t == "Subpixel":
            return subpixel(a, int(l.get("kernel_size", 2)))
        if t == "YOLO":
            rows = [yolo(a[j], ins[1][j], net_hw=(a.shape[1] * 32, a.shape[2] * 32)) for j in range(n)]
            return (np.zeros((n, 1, 1, 1), np.float32), rows)
        raise NotImplementedError("oracle: layer type %s" % t)
