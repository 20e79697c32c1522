"""Images of a batch do not see each other.

The streaming stacks lay the images of a batch side by side in one virtual row per image row, separated by one zero column
(fused_stream.cu, fused_stream_x3.cu), and 128-lane strips run across the image boundaries.  Every image of a batch must
come out as if it had been denoised alone ("same" zero padding at its own borders, module_denoiser.py:39-75 treats the
batch dimension as independent samples).  Kept in the alphabetically last test file: it was written after the round's
GPU budget was spent, so its first run is the driver's.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["f16", "f16x3"])
@pytest.mark.parametrize("n_layers,shape", [(3, (5, 37, 61, 3)), (6, (3, 70, 130, 3)), (2, (7, 9, 127, 3))])
def test_batch_equals_single_images(native_lib, precision, n_layers, shape):
    import blind_image_denoising_b200 as bf
    m = bf.synthetic_model(n_layers, seed=0, precision=precision)
    x = np.random.default_rng(n_layers).integers(0, 256, size=shape, dtype=np.uint8)
    for pad in (False, True):
        yb = m(x, pad_pow2=pad)
        assert yb.shape == x.shape and yb.dtype == np.uint8
        for b in range(shape[0]):
            y1 = m(x[b:b + 1], pad_pow2=pad)
            d = np.abs(yb[b].astype(int) - y1[0].astype(int))
            # the arithmetic per pixel does not depend on the strip position, so this is expected to be 0; the gate is the
            # uint8 parity gate (<= 1 LSB on < 1 % of the values)
            assert d.max() <= 1, (precision, b, pad, int(d.max()))
            assert (d != 0).mean() < 0.01, (precision, b, pad, float((d != 0).mean()))
    m.close()
