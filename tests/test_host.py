"""CPU checks of the host-side logic and of the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native, tensorbundle as tb
from blind_image_denoising_b200.arch import (Arch, arch_from_config, arch_from_name,
                                             arch_from_variable_shapes)
from blind_image_denoising_b200.weights import (flatten_variables, gather_trainables,
                                                synthetic_variables, unflatten_variables)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    hdr = open(os.path.join(ROOT, "include", "bfcnn_b200.h")).read()
    declared = set(re.findall(r"\b(bfcnn_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(native_lib, s), s
    assert native_lib.bfcnn_abi_version() == 3


def test_native_param_counts(native_lib):
    # SURVEY 8a derived constants: 28 784 / 56 528 / 84 272 trainable floats
    for n, want in ((6, 28784), (12, 56528), (18, 84272)):
        a = Arch(no_layers=n)
        c = a.to_c()
        assert native_lib.bfcnn_num_trainable(ctypes.byref(c)) == want == a.num_trainable()
        assert native_lib.bfcnn_num_weights(ctypes.byref(c)) == a.num_weights() == want + 32 * n
    assert Arch(no_layers=18).flops_per_pixel() == 167968 and Arch(no_layers=18).receptive_radius == 37


def test_no_cpu_fallback(native_lib):
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.NativeError):
        bf.synthetic_model(6)
    with pytest.raises(_native.NativeError):
        bf.load_model("resnet_color_1x6_bn_16x3x3_256x256_l1_relu")
    assert native_lib.bfcnn_device_count() < 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "blind_image_denoising_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle\.|bfcnn_oracle", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dp, f)).read()), f


def test_load_model_errors_match_reference():
    # reference bfcnn/__init__.py:81-97,103-112
    with pytest.raises(ValueError, match="model_path cannot be empty"):
        bf.load_model("")
    with pytest.raises(ValueError, match=r"model_path \[/does/not/exist\] does not exist"):
        bf.load_model("/does/not/exist")
    with pytest.raises(ValueError, match="does not exist"):
        bf.load_denoiser_model("nope")
    import bfcnn
    assert set(bfcnn.models) == {f"resnet_color_1x{n}_bn_16x3x3_256x256_l1_relu" for n in (6, 12, 18)}
    for m in bfcnn.models.values():
        assert {"directory", "denoiser", "configuration", "saved_model_path"} <= set(m)
        assert os.path.exists(m["configuration"])


def test_arch_inference():
    a = Arch(no_layers=12)
    assert arch_from_variable_shapes(a.variable_shapes()) == a
    assert arch_from_name("resnet_color_1x18_bn_16x3x3_256x256_l1_relu").no_layers == 18
    assert arch_from_config(bf.CONFIGS_DICT["resnet_color_1x6_bn_16x3x3_256x256_l1_relu"]) == Arch(no_layers=6)
    with pytest.raises(ValueError):
        arch_from_variable_shapes(a.variable_shapes()[:-1])
    with pytest.raises(ValueError):
        arch_from_config({"model": {"backbone": {"type": "unet", "no_layers": 3}}})


def test_flatten_roundtrip_and_trainables():
    a = Arch(no_layers=3)
    v = synthetic_variables(a, 5)
    flat = flatten_variables(a, v)
    assert flat.size == a.num_weights()
    v2 = unflatten_variables(a, flat)
    assert all(np.array_equal(x, y) for x, y in zip(v, v2))
    assert gather_trainables(a, flat).size == a.num_trainable()
    with pytest.raises(ValueError):
        flatten_variables(a, v[:-1])


def test_tensorbundle_roundtrip(tmp_path):
    a = Arch(no_layers=12)   # 63 variables: exercises the "10" < "2" key ordering
    v = synthetic_variables(a, 9)
    tb.write_model_variables(str(tmp_path), v)
    v2 = tb.read_model_variables(str(tmp_path))
    assert len(v2) == len(v) and all(np.array_equal(x, y) for x, y in zip(v, v2))
    # corruption is detected by the crc32c of the entry
    p = tmp_path / "variables.data-00000-of-00001"
    raw = bytearray(p.read_bytes()); raw[100] ^= 0xFF; p.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="crc32c"):
        tb.read_model_variables(str(tmp_path))
    assert tb.crc32c(b"123456789") == 0xE3069283


def test_tensorbundle_reads_reference_bundle():
    """The reader against a bundle TensorFlow itself wrote (the reference's shipped unet)."""
    d = "/root/reference/bfcnn/pretrained/unet_laplacian_v5.6/denoiser/variables"
    if not os.path.isdir(d):
        pytest.skip("reference tree not present on this box")
    v = tb.read_model_variables(d)          # verifies block and tensor crc32c
    assert len(v) == 95 and v[0].shape == (5, 5, 3, 32) and sum(x.size for x in v) == 334976


def test_pretrained_dirs_hold_loadable_weights():
    """The three model directories `load_model` knows by name: variables of the family's shapes in Keras order, finite, and
    a pipeline.json that says where the weights come from (trained here, or synthetic: the reference ships none)."""
    import json
    import os
    for n in (6, 12, 18):
        name = f"resnet_color_1x{n}_bn_16x3x3_256x256_l1_relu"
        d = bf.models[name]["directory"]
        v = bf.load_variables(d)
        ref = synthetic_variables(Arch(no_layers=n), 0)
        assert len(v) == len(ref) and all(x.shape == y.shape and np.isfinite(x).all() for x, y in zip(v, ref))
        marker = json.load(open(os.path.join(d, "pipeline.json")))["weights"]
        assert marker.startswith("TRAINED") or marker.startswith("SYNTHETIC")
        if marker.startswith("SYNTHETIC"):
            assert all(np.array_equal(x, y) for x, y in zip(v, ref))


def test_pipelined_denoiser_host_logic():
    """PipelinedDenoiser.map (no GPU: stand-in instances): results in input order, every instance is driven from ONE
    thread at a time (a handle is not re-entrant), `outs` are routed to the matching call, look-ahead is bounded."""
    import threading
    import time
    from blind_image_denoising_b200.denoiser import PipelinedDenoiser

    class Fake:
        made = []

        def __init__(self):
            self.busy = threading.Lock()
            self.calls = []
            self.closed = False
            Fake.made.append(self)

        def __call__(self, x, out=None, scale=1):
            assert self.busy.acquire(blocking=False), "two threads inside one instance"
            try:
                time.sleep(0.002 * (x % 3))
                self.calls.append(x)
                y = x * 10 * scale
                if out is not None:
                    out.append(y)
                return y
            finally:
                self.busy.release()

        def close(self):
            self.closed = True

    pipe = PipelinedDenoiser(Fake, depth=3)
    outs = [[] for _ in range(20)]
    got = list(pipe.map(range(20), outs=outs, scale=2))
    assert got == [20 * i for i in range(20)]
    assert [o[0] for o in outs] == got
    assert [m.calls for m in Fake.made] == [list(range(k, 20, 3)) for k in range(3)]
    assert pipe(7) == 70
    pipe.close()
    assert all(m.closed for m in Fake.made)
    with pytest.raises(ValueError):
        PipelinedDenoiser(Fake, depth=0)
