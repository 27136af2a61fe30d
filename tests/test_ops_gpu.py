"""GPU parity tests, operator level: every C-ABI operator launch vs the oracle on the same seeded inputs.

Mirrors the reference's op tests (demo/test/unittest/*Test.cpp): same PRNG + seed (prng.h, 7767517), same input
constructions (all-ones x RandomMat U[-1.2,1.2]; pooling sentinels), same grids, same comparator
(|a-b|<=eps or rel<eps, testutil.cpp:351-361) — but with eps = 1e-3 (north-star tolerance) instead of the
reference's 0.01.
"""
import zlib

import numpy as np
import pytest

from oracle import oracle
from shadernn_b200 import core

pytestmark = pytest.mark.gpu

EPS = 1e-3  # north_star: within 1e-3 relative fp32 per layer


def assert_close(got, want, eps=EPS, what=""):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    bad = oracle.compare(got, want, eps)
    if bad:
        d = np.abs(got - want)
        i = np.unravel_index(np.argmax(d), d.shape)
        raise AssertionError("%s: %d/%d elements differ (eps %g); worst at %s: got %r want %r" % (what, bad, got.size, eps, i, got[i], want[i]))
    # and the typical error must be far inside the tolerance (split-bf16 storage keeps ~17 bits)
    denom = np.maximum(np.abs(want), 1.0)
    assert float(np.max(np.abs(got - want) / denom)) < 2e-4, what


def conv_case(ctx, n, h, w, ic, oc, k, s, padding, act="", alpha=0.0, bias=True, bn=False, mode="constant", ones=False, algo="simt", residual=False, seed=0):
    rng = np.random.default_rng(seed)
    if ones:
        oracle.srand(7767517)
        x = np.ones((n, h, w, ic), np.float32)
        wt = oracle.random_mat((oc, ic, k, k))
    else:
        x = rng.uniform(-1, 1, (n, h, w, ic)).astype(np.float32)
        wt = (rng.standard_normal((oc, ic, k, k)) * np.sqrt(2.0 / (k * k * ic))).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, oc).astype(np.float32) if bias else None
    bnd = None
    if bn:
        bnd = {"gamma": rng.uniform(0.5, 1.5, oc), "beta": rng.uniform(-0.1, 0.1, oc), "mean": rng.uniform(-0.1, 0.1, oc), "var": rng.uniform(0.5, 1.5, oc)}
        if ones:
            bnd = {"gamma": np.ones(oc), "beta": np.zeros(oc), "mean": np.zeros(oc), "var": np.ones(oc)}
    if isinstance(padding, str):
        o = oracle.same_padding(k, padding == "same")
    else:
        o = list(padding)
    oh, ow = oracle.conv_out_dim(h, k, s, o[0], o[1]), oracle.conv_out_dim(w, k, s, o[0], o[1])
    px, py = (0, 0) if k == 1 else (o[0], o[2])
    want = oracle.conv2d(x, wt, b, bnd, s, px, py, mode, act if not residual else "", alpha, (oh, ow))
    res = None
    if residual:
        res = rng.uniform(-1, 1, want.shape).astype(np.float32)
        want = oracle.add(want, res, act, alpha)
    got = core.conv2d(ctx, x, wt, b, bnd, s, px, py, mode, act, alpha, (oh, ow), residual=res, algo=algo)
    return got, want


# the reference's convolution grid: W=H=8, C=128, OC=1, k=1 by default; CLI sweeps -K 1..5, -S 1..2 (convolutionTest.cpp:419-424)
@pytest.mark.parametrize("k,s", [(1, 1), (1, 2), (3, 1), (3, 2), (5, 1), (2, 1), (4, 2), (7, 2)])
def test_conv2d_reference_grid_ones_input(ctx, k, s):
    got, want = conv_case(ctx, 1, 8, 8, 128, 1, k, s, "same", bias=False, bn=True, ones=True)
    assert_close(got, want, what="conv ones k%d s%d" % (k, s))


@pytest.mark.parametrize("shape", [
    # n, h, w, ic, oc, k, s, padding
    (2, 14, 14, 64, 64, 3, 1, "same"),
    (2, 15, 13, 3, 32, 3, 2, "same"),      # C_in = 3 stem, odd dims
    (1, 16, 16, 3, 64, 7, 2, "same"),      # resnet stem
    (3, 9, 9, 24, 144, 1, 1, "valid"),     # mobilenet expand
    (2, 9, 9, 144, 24, 1, 1, "valid"),     # mobilenet project
    (1, 8, 8, 256, 255, 1, 1, "valid"),    # yolo head: OC not a multiple of 8
    (1, 12, 12, 1, 16, 5, 1, "same"),      # espcn first layer: single input channel
    (1, 12, 12, 16, 4, 3, 1, "same"),      # espcn last layer
    (2, 7, 7, 512, 512, 3, 1, "same"),
    (1, 10, 10, 20, 12, 3, 1, "valid"),    # IC not a multiple of 8
    (1, 6, 6, 32, 32, 3, 1, (1, 0, 1, 0)), # asymmetric numeric padding [T,B,L,R]
])
@pytest.mark.parametrize("act", ["", "relu", "relu6", "leakyRelu", "tanh", "sigmoid", "SiLU"])
def test_conv2d_shapes_and_epilogues(ctx, shape, act):
    n, h, w, ic, oc, k, s, padding = shape
    got, want = conv_case(ctx, n, h, w, ic, oc, k, s, padding, act=act, alpha=0.1, bias=True, bn=True, seed=zlib.crc32(repr((shape, act)).encode()) % 1000)
    assert_close(got, want, what="conv %s %s" % (shape, act))


@pytest.mark.parametrize("mode", ["constant", "replicate", "reflect"])
def test_conv2d_padding_modes(ctx, mode):
    got, want = conv_case(ctx, 2, 11, 9, 16, 24, 3, 1, (1, 1, 1, 1), act="relu", mode=mode)
    assert_close(got, want, what="conv pad " + mode)
    got, want = conv_case(ctx, 1, 20, 20, 3, 8, 9, 1, (4, 4, 4, 4), act="", mode=mode)  # candy 9x9
    assert_close(got, want, what="conv9 pad " + mode)


def test_conv2d_fused_residual(ctx):
    got, want = conv_case(ctx, 2, 14, 14, 64, 64, 3, 1, "same", act="relu", bn=True, residual=True)
    assert_close(got, want, what="conv+add+relu")


def test_conv2d_empty_and_ragged_edges(ctx):
    # M not a multiple of the 64-pixel tile, OC not a multiple of the 64-channel tile
    got, want = conv_case(ctx, 1, 5, 7, 8, 72, 3, 1, "same")
    assert_close(got, want, what="ragged tiles")
    got, want = conv_case(ctx, 1, 1, 1, 1280, 1000, 1, 1, "valid")  # 1x1 spatial (classifier head as conv)
    assert_close(got, want, what="1x1 spatial")


# the reference's depthwise grid (9,9,8,8,k1..) (depthwiseConv2DTest.cpp:316)
@pytest.mark.parametrize("h,w,c,k,s,padding", [(9, 9, 8, 1, 1, "same"), (9, 9, 8, 3, 1, "same"), (9, 9, 8, 3, 2, "same"), (14, 14, 96, 3, 1, "same"),
                                              (15, 15, 144, 3, 2, "valid"), (10, 12, 20, 5, 1, "same"), (8, 8, 32, 3, 2, (0, 1, 0, 1))])
@pytest.mark.parametrize("act", ["", "relu6"])
def test_depthwise(ctx, h, w, c, k, s, padding, act):
    rng = np.random.default_rng(k * 100 + c)
    x = rng.uniform(-1, 1, (2, h, w, c)).astype(np.float32)
    wt = (rng.standard_normal((c, k, k)) * np.sqrt(2.0 / (k * k))).astype(np.float32)
    bn = {"gamma": rng.uniform(0.5, 1.5, c), "beta": rng.uniform(-0.1, 0.1, c), "mean": rng.uniform(-0.1, 0.1, c), "var": rng.uniform(0.5, 1.5, c)}
    o = oracle.same_padding(k, padding == "same") if isinstance(padding, str) else list(padding)
    ow = oracle.depthwise_out_dim(w, k, s, o[0], o[2])
    oh = oracle.depthwise_out_dim(h, k, s, o[1], o[3])
    want = oracle.depthwise(x, wt, None, bn, s, o[0], o[2], act, 0.0, (oh, ow))
    got = core.depthwise(ctx, x, wt, None, bn, s, o[0], o[2], act, 0.0, (oh, ow))
    assert_close(got, want, what="depthwise")


# the reference's pooling grid (9,9,4,avg,k2,s3,'same') with sentinels 100/200 (poolingTest.cpp:42-44,104)
def test_pooling_reference_grid(ctx):
    x = np.full((1, 9, 9, 4), 100.0, np.float32)
    x[0, ::3, ::3, :] = 200.0
    oh = oracle.pool_out_dim(9, 2, 3, False)
    for avg in (True, False):
        want = oracle.pool2d(x, 2, 3, avg, (oh, oh))
        got = core.pool2d(ctx, x, 2, 3, avg, (oh, oh))
        assert np.array_equal(got, want)  # sentinels are exactly representable


@pytest.mark.parametrize("h,w,c,k,s,valid", [(112, 112, 64, 3, 2, False), (13, 13, 512, 2, 1, False), (26, 26, 256, 2, 2, True), (7, 7, 1280, 7, 1, True), (7, 7, 1280, 7, 7, True),
                                             (5, 5, 3, 3, 2, False), (9, 7, 12, 2, 2, True)])
@pytest.mark.parametrize("avg", [False, True])
def test_pooling(ctx, h, w, c, k, s, valid, avg):
    x = np.random.default_rng(h + c).uniform(-2, 2, (2, h, w, c)).astype(np.float32)
    oh, ow = oracle.pool_out_dim(h, k, s, valid), oracle.pool_out_dim(w, k, s, valid)
    want = oracle.pool2d(x, k, s, avg, (oh, ow))
    got = core.pool2d(ctx, x, k, s, avg, (oh, ow))
    assert_close(got, want, what="pool")


# add 4x4x1 (binaryOpTest.cpp:116), activation 16x16x4 SiLU (activationTest.cpp:116)
def test_add_and_activation(ctx):
    rng = np.random.default_rng(3)
    a, b = rng.uniform(-3, 3, (1, 4, 4, 1)).astype(np.float32), rng.uniform(-3, 3, (1, 4, 4, 1)).astype(np.float32)
    for act in ["", "relu", "relu6", "tanh", "sigmoid", "leaky_relu", "SiLU"]:
        assert_close(core.add(ctx, a, b, act, 0.3), oracle.add(a, b, act, 0.3), what="add " + act)
    x = rng.uniform(-4, 8, (2, 16, 16, 4)).astype(np.float32)
    for act in ["relu", "relu6", "tanh", "sigmoid", "leaky_relu", "SiLU"]:
        assert_close(core.activation(ctx, x, act, 0.2), oracle.activation(x, act, 0.2), what="activation " + act)
    big_a, big_b = rng.uniform(-3, 3, (3, 28, 28, 130)).astype(np.float32), rng.uniform(-3, 3, (3, 28, 28, 130)).astype(np.float32)
    assert_close(core.add(ctx, big_a, big_b, "relu"), oracle.add(big_a, big_b, "relu"), what="add big")


def test_batchnorm_and_instancenorm(ctx):
    rng = np.random.default_rng(4)
    x = rng.uniform(-2, 2, (2, 9, 9, 20)).astype(np.float32)
    bn = {"gamma": rng.uniform(0.5, 1.5, 20), "beta": rng.uniform(-1, 1, 20), "mean": rng.uniform(-1, 1, 20), "var": rng.uniform(0.2, 2, 20)}
    assert_close(core.batchnorm(ctx, x, bn, "relu"), oracle.batchnorm(x, bn, "relu"), what="batchnorm")
    ones = {"gamma": np.ones(20), "beta": np.zeros(20), "mean": np.zeros(20), "var": np.ones(20)}
    assert_close(core.batchnorm(ctx, np.ones_like(x), ones), np.full_like(x, 1 / np.sqrt(1.001)), what="bn factor")  # batchNormTest construction
    y = rng.standard_normal((2, 33, 31, 12)).astype(np.float32) * 3 + 1
    g, b = rng.uniform(0.5, 1.5, 12).astype(np.float32), rng.uniform(-0.5, 0.5, 12).astype(np.float32)
    assert_close(core.instancenorm(ctx, y, g, b, "relu"), oracle.instancenorm(y, g, b, "relu"), what="instancenorm")


# dense 11 -> 5 (denseTest.cpp:111), flatten 1x1x23 (flattenTest.cpp:88)
def test_dense_flatten_softmax_argmax(ctx):
    rng = np.random.default_rng(5)
    for n_in, n_out in [(11, 5), (512, 10), (1280, 1000)]:
        x = rng.uniform(-1, 1, (3, 1, 1, n_in)).astype(np.float32)
        k = rng.uniform(-1.2, 1.2, (n_out, n_in)).astype(np.float32) / np.sqrt(n_in)
        b = rng.uniform(-0.5, 0.5, n_out).astype(np.float32)
        for act in ["", "relu", "softmax", "sigmoid", "tanh", "leaky_relu"]:
            assert_close(core.dense(ctx, x, k, b, act, 0.3), oracle.dense(x, k, b, act, 0.3), what="dense %s" % act)
    f = rng.uniform(-1, 1, (2, 1, 1, 23)).astype(np.float32)
    assert_close(core.flatten(ctx, f), oracle.flatten(f), what="flatten 1x1x23")
    f2 = rng.uniform(-1, 1, (2, 3, 4, 5)).astype(np.float32)
    assert_close(core.flatten(ctx, f2), oracle.flatten(f2), what="flatten HWC")
    # dense on a non-1x1 tensor consumes it in HWC order (CPU flatten, cpulayer.h:94-115)
    k2 = rng.uniform(-1, 1, (7, 60)).astype(np.float32)
    assert_close(core.dense(ctx, f2, k2, None, ""), oracle.dense(oracle.flatten(f2), k2, None, ""), what="dense on HWC")
    logits = rng.uniform(-5, 5, (4, 1, 1, 10)).astype(np.float32)
    assert_close(core.softmax(ctx, logits), oracle.softmax(logits), what="softmax")
    assert np.array_equal(core.argmax1(ctx, logits), oracle.argmax1(logits))  # 1-based index, bit-exact
    ties = np.zeros((2, 1, 1, 40), np.float32)
    ties[0, 0, 0, [7, 33]] = 2.0
    assert list(core.argmax1(ctx, ties)) == [8, 1]  # first maximum wins


def test_layout_ops(ctx):
    rng = np.random.default_rng(6)
    a = rng.uniform(-1, 1, (2, 13, 13, 128)).astype(np.float32)
    b = rng.uniform(-1, 1, (2, 13, 13, 255)).astype(np.float32)
    assert_close(core.concat(ctx, a, b), oracle.concat(a, b), what="concat")
    assert_close(core.upsample(ctx, a, 2, False), oracle.upsample(a, 2, False), what="upsample nearest")
    small = rng.uniform(-1, 1, (1, 5, 6, 9)).astype(np.float32)
    assert_close(core.upsample(ctx, small, 2, True), oracle.upsample(small, 2, True), what="upsample bilinear")
    for mode in ["constant", "replicate", "reflect"]:
        assert_close(core.pad(ctx, small, 2, 1, (5 + 1 + 3, 6 + 2 + 2), mode), oracle.pad(small, 2, 1, (9, 10), mode), what="pad " + mode)
    z = rng.uniform(-2, 2, (2, 7, 9, 4)).astype(np.float32)
    assert_close(core.subpixel(ctx, z, 2), oracle.subpixel(z, 2), what="subpixel")


def test_tensor_layouts_and_dump(ctx, tmp_path):
    rng = np.random.default_rng(7)
    x = rng.uniform(-100, 100, (2, 5, 7, 6)).astype(np.float32)
    t = core.ImageTexture.from_numpy(ctx, x)
    back = t.download()
    # split-bf16 storage: relative error <= 2^-16
    assert np.max(np.abs(back - x) / np.maximum(np.abs(x), 1e-30)) < 2 ** -15
    c4 = t.download_c4hw4()
    assert c4.shape == (2, 2, 5, 7, 4)
    assert np.array_equal(c4[:, 1, :, :, 1], back[..., 5]) and np.all(c4[:, 1, :, :, 2:] == 0)
    t2 = core.ImageTexture(ctx, 2, 5, 7, 6)
    t2.upload_c4hw4(c4)
    assert np.array_equal(t2.download(), back)  # values already representable: exact round trip
    p = str(tmp_path / "t.dump")
    t.dump(p)
    raw = open(p + ".n1", "rb").read()
    assert raw[:32].rstrip(b"\0") == b"7 5 2 6"  # "W H D C" (image.cpp:216-245)
    assert np.array_equal(np.frombuffer(raw[32:], np.float32).reshape(2, 5, 7, 4), c4[1])


def test_error_paths(ctx):
    from shadernn_b200._lib import SnnbError
    with pytest.raises(SnnbError):
        core.ImageTexture(ctx, 0, 1, 1, 1)
    a = np.zeros((1, 4, 4, 8), np.float32)
    with pytest.raises(SnnbError):
        core.add(ctx, a, np.zeros((1, 4, 4, 16), np.float32))
    with pytest.raises(SnnbError):  # tcgen05 forced on a shape it cannot take
        core.conv2d(ctx, np.zeros((1, 4, 4, 3), np.float32), np.zeros((8, 3, 7, 7), np.float32), out_hw=(4, 4), pad_x=3, pad_y=3, pad_mode="reflect",
                    algo="tcgen05")


def test_backend_level_launch_capture(ctx):
    # INTEGRATION.md depth B: a DeviceBackend built on the per-operator calls records its stage loop once
    # (snnb_graph_capture_begin / _end) and replays it. conv3x3(+relu) -> add(+relu) -> maxpool, replayed on new input.
    import ctypes as C
    from shadernn_b200._lib import lib, check
    rng = np.random.default_rng(21)
    n, h, w, ic, oc = 2, 28, 28, 64, 64
    wt = (rng.standard_normal((oc, ic, 3, 3)) * np.sqrt(2.0 / (9 * ic))).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, oc).astype(np.float32)
    d = core.conv_desc(ic, oc, 3, 1, 1, 1, "constant", "relu", 0.0, "auto")
    wh = C.c_void_p()
    check(lib().snnb_weights_pack_conv2d(ctx.h, C.byref(d), wt.ctypes.data_as(C.c_void_p), bias.ctypes.data_as(C.c_void_p), None, None, None, None, C.byref(wh)))
    wobj = core.Weights(wh)
    tin, tmid, tsum, tout = (core.ImageTexture(ctx, n, h, w, c) for c in (ic, oc, oc, oc))
    tpool = core.ImageTexture(ctx, n, 14, 14, oc)

    def stage_loop():
        check(lib().snnb_conv2d_launch(ctx.h, C.byref(d), wobj.h, tin.h, None, tmid.h))
        check(lib().snnb_add_launch(ctx.h, core.ACT["relu"], 0.0, tmid.h, tin.h, tsum.h))
        check(lib().snnb_maxpool_launch(ctx.h, 2, 2, tsum.h, tpool.h))

    def expected(x):
        y = oracle.conv2d(x, wt, bias=bias, stride=1, pad_x=1, pad_y=1, activation="relu", out_hw=(h, w))
        return oracle.pool2d(oracle.add(y, x, activation="relu"), 2, 2, False, (14, 14))

    x0 = rng.uniform(-1, 1, (n, h, w, ic)).astype(np.float32)
    tin.upload(x0)
    stage_loop()  # eager pass first: lazily created resources must exist before capture
    ctx.sync()
    check(lib().snnb_graph_capture_begin(ctx.h))
    stage_loop()
    g = C.c_void_p()
    check(lib().snnb_graph_capture_end(ctx.h, C.byref(g)))
    try:
        for seed in (1, 2):
            x = np.random.default_rng(seed).uniform(-1, 1, (n, h, w, ic)).astype(np.float32)
            tin.upload(x)
            check(lib().snnb_graph_launch(g))
            got = tpool.download()
            assert_close(got, expected(x))
    finally:
        lib().snnb_graph_destroy(g)
