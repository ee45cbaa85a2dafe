"""Observation-frame preprocessing: the numpy restatement of Pillow's resampler is pinned on Pillow itself,
the product's coefficient tables equal the restatement's, the host orchestration (DeviceImageProcessor over the torch op
re-statement) equals the restated SiglipImageProcessor, and -- on the GPU -- the CUDA kernels equal it bit for bit."""
import numpy as np
import pytest
import torch

from oracle import image_ref as IR
from oracle.ops_ref import RefOps
from spatialvla_b200.image_processing import DeviceImageProcessor, resample_tables, value_lut

SIZES = ((480, 640), (256, 256), (224, 224), (128, 160), (300, 224), (224, 500), (720, 1280))


def _img(rng, h, w):
    base = rng.random((h // 8 + 2, w // 8 + 2, 3))
    im = np.kron(base, np.ones((8, 8, 1)))[:h, :w] * 200 + rng.random((h, w, 3)) * 55       # structure + noise
    return im.astype(np.uint8)


def test_restatement_equals_pillow():
    from PIL import Image
    rng = np.random.default_rng(0)
    for h, w in SIZES:
        im = _img(rng, h, w)
        ref = np.asarray(Image.fromarray(im).resize((224, 224), resample=Image.BICUBIC))
        assert np.array_equal(IR.resize_u8_ref(im), ref), (h, w)


def test_product_tables_equal_restatement_and_host_logic():
    rng = np.random.default_rng(1)
    for n in (640, 480, 256, 160, 128, 1280, 225):
        b, k, ks = resample_tables(n, 224)
        rb, rk = IR.resample_coeffs(n, 224)
        assert np.array_equal(b, rb) and np.array_equal(k, rk) and ks == rk.shape[1]
    for norm in (False, True):
        proc = DeviceImageProcessor(RefOps(), do_normalize=norm)
        for h, w in ((480, 640), (224, 224), (300, 224), (224, 500)):
            ims = [_img(rng, h, w) for _ in range(2)]
            got = proc(ims).numpy()
            assert np.array_equal(got, IR.preprocess_ref(ims, do_normalize=norm)), (h, w, norm)
    assert value_lut().shape == (3, 256) and value_lut()[0, 255] == np.float32(255 * (1 / 255))
    with pytest.raises(ValueError):
        proc(np.zeros((1, 8, 8, 3), dtype=np.float32))


def test_processor_uses_the_device_path_for_uint8_frames():
    from test_processor_tokenizer_host import ACTION_CONFIG, INTR, STATS
    from fakes import FakeImageProcessor, FakeTokenizer
    from spatialvla_b200 import SpatialVLAProcessor
    p = SpatialVLAProcessor(FakeImageProcessor(), FakeTokenizer(), statistics=STATS, intrinsic_config=INTR, action_config=ACTION_CONFIG,
                            action_chunk_size=4).enable_device_images(ops=RefOps())
    rng = np.random.default_rng(2)
    ims = [_img(rng, 480, 640), _img(rng, 480, 640)]
    out = p(images=ims, text=["pick up the cup", "open drawer"], unnorm_key="bridge")
    assert out["pixel_values"].shape == (2, 3, 224, 224)
    assert np.array_equal(out["pixel_values"].numpy(), IR.preprocess_ref(ims))
    out = p(images=ims, text=["a", "b"], do_normalize=True)
    assert np.array_equal(out["pixel_values"].numpy(), IR.preprocess_ref(ims, do_normalize=True))


@pytest.mark.gpu
def test_cuda_kernels_equal_pillow_restatement(cuda_device):
    from spatialvla_b200.ops import CudaOps
    ops = CudaOps(cuda_device)
    rng = np.random.default_rng(3)
    n0 = ops.launch_count()
    for norm in (False, True):
        proc = DeviceImageProcessor(ops, do_normalize=norm)
        for h, w in SIZES:
            ims = [_img(rng, h, w) for _ in range(3)]
            got = proc(ims).cpu().numpy()
            ref = IR.preprocess_ref(ims, do_normalize=norm)
            assert np.array_equal(got, ref), (h, w, norm, float(np.abs(got - ref).max()))
    assert ops.launch_count() > n0
