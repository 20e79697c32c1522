"""Latency of ONE drop-in call from host memory (numpy uint8 in, numpy uint8 out), the way a user of the reference calls a
model: python tools/latency_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bfcnn
rng = np.random.default_rng(0)
for name in ("resnet_color_1x6_bn_16x3x3_256x256_l1_relu", "resnet_color_1x18_bn_16x3x3_256x256_l1_relu"):
    for prec in ("f16x3", "f16"):
        m = bfcnn.load_model(name, precision=prec)
        for shape in ((1, 256, 256, 3), (1, 512, 512, 3), (1, 1080, 1920, 3), (1, 2160, 3840, 3)):
            x = rng.integers(0, 256, size=shape, dtype=np.uint8)
            for _ in range(3):
                m(x)
            t = []
            for _ in range(20):
                t0 = time.perf_counter(); m(x); t.append(time.perf_counter() - t0)
            t.sort()
            print(f"{name[13:17]:5s} {prec:6s} {shape[1]:5d}x{shape[2]:<5d} median {t[10] * 1e3:8.3f} ms  min {t[0] * 1e3:8.3f} ms  {shape[1] * shape[2] / 1e6 / t[10]:8.1f} MP/s", flush=True)
        m.close()
