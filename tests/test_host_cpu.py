"""CPU suite (-m "not gpu"): host logic and the drop-in boundary.

 * libsnn_b200.so loads and exports every symbol include/snnb.h declares (and _lib.py binds exactly that set);
 * without a GPU the library fails LOUDLY (no CPU fallback anywhere in the product path);
 * the model writers and the two independent readers (Python oracle walker here; the C++ ModelParser is exercised
   on the GPU box) agree on the format: embedded-JSON and split .json+.bin variants decode to identical weights;
 * the multi-GPU plumbing (shard ranges, the single weight broadcast) under gloo with world_size 2.
"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from shadernn_b200 import _lib, modelzoo, parallel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "snnb.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(snnb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) > 50
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.lib()  # raises AttributeError for any missing export
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.snnb_version() == 100


def test_no_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from shadernn_b200.core import GpuContext
    with pytest.raises(_lib.SnnbError) as e:
        GpuContext(0)
    assert "no CPU fallback" in str(e.value) or "no CUDA device" in str(e.value)


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libsnn_b200.so"))
    with pytest.raises(_lib.SnnbError) as e:
        _lib.lib()
    assert "no CPU or PyTorch fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under shadernn_b200/ may reference it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "shadernn_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "snn_oracle" not in text and "from oracle" not in text and "import oracle" not in text, os.path.join(dirpath, f)


@pytest.mark.parametrize("name,hw", [("espcn", (16, 16)), ("resnet18", (32, 32)), ("mobilenetv2", (32, 32)), ("yolov3tiny", (64, 64)), ("candy", (16, 16))])
def test_writer_variants_decode_identically(built, tmp_path, name, hw):
    from oracle import oracle
    fn = modelzoo.MODELS[name][0]
    layers = fn(hw)
    pa = modelzoo.write_model(layers, str(tmp_path / "a.json"), split=False)
    pb = modelzoo.write_model(layers, str(tmp_path / "b_layers.json"), split=True)
    assert os.path.exists(str(tmp_path / "b_weights.bin"))
    ma, mb = oracle.Model(pa), oracle.Model(pb)
    assert ma.count == mb.count == len(layers)
    for wa, wb in zip(ma.w, mb.w):
        assert wa.keys() == wb.keys()
        for k in wa:
            if isinstance(wa[k], dict):
                for kk in wa[k]:
                    assert np.array_equal(wa[k][kk], wb[k][kk])
            else:
                assert np.array_equal(wa[k], wb[k]), k
    x = modelzoo.synthetic_input(name, 1, hw)
    ya, yb = ma.run(x), mb.run(x)
    assert np.array_equal(ya, yb)


def test_model_graphs_have_the_reference_layer_counts():
    # SNN layer counts after the converter's BN/activation fusion: ResNet-18 = 33 (resnet18Test.cpp:85-198)
    assert len(modelzoo.resnet18()) == 33
    r = modelzoo.resnet18()
    assert sum(1 for l in r if l["type"] == "Conv2D") == 20 and sum(1 for l in r if l["type"] == "Add") == 8
    m = modelzoo.mobilenetv2()
    assert sum(1 for l in m if l["type"] == "DepthwiseConv2D") == 17 and sum(1 for l in m if l["type"] == "Add") == 10
    assert sum(1 for l in m if l["type"] == "Conv2D") == 36  # 36 1x1/3x3 conv layers incl. the stem
    y = modelzoo.yolov3_tiny()
    assert sum(1 for l in y if l["type"] == "Conv2D") == 13 and sum(1 for l in y if l["type"] == "MaxPooling2D") == 6
    c = modelzoo.candy()
    assert sum(1 for l in c if l["type"] == "InstanceNormalization") == 15 and sum(1 for l in c if l["type"] == "Conv2D" and l.get("mode") == "reflect") == 16
    assert len(modelzoo.espcn()) == 5


def test_shard_ranges_cover_and_partition():
    for total in (0, 1, 7, 8, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for s, c in spans:
                assert s == pos
                pos += c
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)
    assert parallel.shard_range(64, 8, 3) == (24, 8)  # MobileNetV2 batch 64 over 8 GPUs


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch
from shadernn_b200 import parallel
rank, local, world = parallel.init_distributed(backend="gloo")
assert world == 2
# stand-in for the packed weight arena: rank 0 holds the real bytes, rank 1 garbage
arena = torch.arange(4096, dtype=torch.uint8) if rank == 0 else torch.full((4096,), 7, dtype=torch.uint8)
parallel.broadcast_buffer(arena, src=0)
assert torch.equal(arena, torch.arange(4096, dtype=torch.uint8)), rank
start, count = parallel.shard_range(9, world, rank)
ms = parallel.max_over_ranks(10.0 + rank)
assert ms == 11.0
parallel.barrier()
print("OK", rank, start, count)
"""


def test_two_rank_gloo_weight_broadcast_and_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2")
    procs = []
    for rank in range(2):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            p.kill()
            raise
        outs.append(o)
        assert p.returncode == 0, o
    assert "OK 0 0 5" in outs[0] and "OK 1 5 4" in outs[1]
