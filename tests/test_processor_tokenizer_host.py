"""CPU: host-side API surfaces (processor string building, intrinsics, decode/un-normalise plumbing, tokenizer
configuration). The grid arithmetic itself needs the GPU (tests/test_kernels_gpu.py::tokenizer_case)."""
import numpy as np
import pytest
import torch

from fakes import FakeImageProcessor, FakeTokenizer

ACTION_CONFIG = {"num_bins": {"translation": {"theta_bins": 16, "phi_bins": 32, "r_bins": 8},
                              "rotation": {"roll_bins": 16, "pitch_bins": 16, "yaw_bins": 16}, "gripper": 2, "total": 8194},
                 "use_spherical": True}
INTR = {"default": {"intrinsic": [[623.588, 0, 319.501], [0, 623.588, 239.545], [0, 0, 1]], "height": 480, "width": 640}}
STATS = {"bridge": {"action": {"q01": [-0.1] * 7, "q99": [0.2] * 7, "mask": [True] * 6 + [False]}}}


def make_processor(**kw):
    from spatialvla_b200 import SpatialVLAProcessor
    return SpatialVLAProcessor(FakeImageProcessor(), FakeTokenizer(), statistics=STATS, intrinsic_config=INTR,
                               action_config=ACTION_CONFIG, action_chunk_size=4, **kw)


def test_processor_prompt_layout_and_intrinsics():
    p = make_processor()
    tk = p.action_tokenizer
    assert tk.vocab_size == 8194 and tk.action_token_begin_idx == 257152 + 0 or tk.action_token_begin_idx > 0
    img = (np.random.default_rng(0).random((224, 224, 3)) * 255).astype(np.uint8)
    out = p(images=[img, img], text=["pick up the cup", "open drawer"], unnorm_key="bridge")
    ids = out["input_ids"]
    assert out["pixel_values"].shape == (2, 3, 224, 224)
    assert (ids[:, :256] == p.image_token_id).all() and (ids[:, 256] == 2).all()
    row0 = ids[0][out["attention_mask"][0] == 1]
    assert row0[-1] == 108                                   # trailing newline token
    K = out["intrinsic"]
    assert K.shape == (3, 3)
    np.testing.assert_allclose(K.numpy(), [[218.2558, 0, 111.82535], [0, 291.0077, 111.78767], [0, 0, 1]], rtol=1e-5)
    with pytest.raises(ValueError):
        p(images=[img], text=["a", "b"])
    with pytest.raises(ValueError):
        p(images=None, text="a")
    moved = out.to(torch.float16)
    assert moved["pixel_values"].dtype == torch.float16 and moved["input_ids"].dtype == torch.int64


def test_tokenizer_sub_ranges_and_edges():
    from spatialvla_b200 import SpatialActionTokenizer
    t = SpatialActionTokenizer(FakeTokenizer(257153), ACTION_CONFIG["num_bins"], gs_params=None)
    b = t.action_token_begin_idx
    assert (t.translation_tokenizer.token_start_idx, t.translation_tokenizer.token_end_idx) == (b, b + 4095)
    assert (t.rotation_tokenizer.token_start_idx, t.rotation_tokenizer.token_end_idx) == (b + 4096, b + 8191)
    assert (t.gripper_tokenizer.token_start_idx, t.gripper_tokenizer.token_end_idx) == (b + 8192, b + 8193)
    assert t._edges.shape == (17 + 33 + 9 + 17 * 3,)
    assert t.token_array[4096] == "<ACTION04096>"
    with pytest.raises(ValueError):
        SpatialActionTokenizer(FakeTokenizer(), ACTION_CONFIG["num_bins"],
                               bin_policy={"translation": {"theta_bins": [0, 1], "phi_bins": [0, 1], "r_bins": [0, 1]},
                                           "rotation": {"roll_bins": [0, 1], "pitch_bins": [0, 1], "yaw_bins": [0, 1]}})


def test_unnormalise_formula():
    p = make_processor()
    n = np.array([[0.0, 1.0, -1.0, 0.5, 0.0, 0.0, 1.0]])
    a = p._unnormalize(n, "bridge")
    np.testing.assert_allclose(a[0, :3], [0.05, 0.2, -0.1])
    assert a[0, 6] == 1.0                                     # masked-out gripper passes through
