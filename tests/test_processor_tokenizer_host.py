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


def test_training_call_routes_kwargs_like_the_reference_dataset():
    """data/dataset.py:134-142 calls the processor with image options (do_normalize=False) and tokenizer options (max_length,
    truncation, padding) in one kwargs bag; HF's _merge_kwargs routes them, and max_length is raised by the 256 image tokens
    (model/processing_spatialvla.py:113-117,176-178)."""
    seen = {}

    class RecImage(FakeImageProcessor):
        def __call__(self, images, return_tensors="pt", **kw):
            seen["image"] = dict(kw)
            return super().__call__(images, return_tensors=return_tensors)

    class RecTok(FakeTokenizer):
        def __call__(self, strings, **kw):
            seen["text"] = {k: v for k, v in kw.items() if k not in ("text_pair", "return_token_type_ids", "return_tensors")}
            return super().__call__(strings, **kw)

    from spatialvla_b200 import SpatialVLAProcessor
    p = SpatialVLAProcessor(RecImage(), RecTok(), statistics=STATS, intrinsic_config=INTR, action_config=ACTION_CONFIG, action_chunk_size=4)
    from oracle import tokenizer_ref as T
    tk, nb = p.action_tokenizer, ACTION_CONFIG["num_bins"]           # grid arithmetic needs the GPU; the host test injects the oracle
    tk.encode_local_ids = lambda a: T.encode(np.asarray(a, dtype=np.float64).reshape(-1, 7), tk.bin_policy, nb).astype(np.int32)
    img = (np.random.default_rng(0).random((224, 224, 3)) * 255).astype(np.uint8)
    acts = np.random.default_rng(1).uniform(-1, 1, (4, 7))
    out = p(text="pick up the cup", images=[img], suffix_actions=acts, return_tensors="pt", padding=False, max_length=64,
            truncation=True, do_normalize=False, not_an_option=1)
    assert seen["image"] == {"do_normalize": False}
    assert seen["text"] == {"padding": False, "max_length": 64 + 256, "truncation": True}
    assert out["labels"].shape == out["input_ids"].shape and "token_type_ids" in out


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


def test_training_sample_chain_processor_to_labelled_forward():
    """The training-sample path of the reference end to end on the host: SpatialVLAProcessor(suffix_actions=...) builds prefix +
    action tokens + EOS with token types and labels (model/processing_spatialvla.py:103-192, train/monkey_patch.py:21-75), the
    model's forward(labels=...) consumes it (prefix-LM mask) and action_metrics de-tokenises the predictions.  The grid kernels
    behind the tokenizer need a GPU, so their oracle restatement is patched in; the model runs on the torch op re-statements."""
    from oracle import model_ref as R
    from oracle import tokenizer_ref as T
    from oracle.ops_ref import RefOps
    from spatialvla_b200 import SpatialVLAProcessor
    from spatialvla_b200.configs import get_config_dict
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    from spatialvla_b200.weights import synth_state_dict
    tok = FakeTokenizer(base=1010)                    # text ids of the fake tokenizer stay below 1003
    p = SpatialVLAProcessor(FakeImageProcessor(), tok, statistics=STATS, intrinsic_config=INTR, action_config=ACTION_CONFIG,
                            action_chunk_size=2)
    tk = p.action_tokenizer
    nb = ACTION_CONFIG["num_bins"]
    tk.encode_local_ids = lambda a: T.encode(np.asarray(a, dtype=np.float64).reshape(-1, 7), tk.bin_policy, nb).astype(np.int32)
    tk.decode_token_ids_to_actions = lambda ids: T.decode(np.asarray(ids) - tk.action_token_begin_idx, tk.bin_policy, nb)
    # the tiny model config follows the tokenizer (like a real checkpoint's config.json does): image / action token ids, vocabulary
    cfg = get_config_dict("tiny")
    lo = tk.action_token_begin_idx
    cfg["image_token_index"], cfg["action_token_begin_idx"] = p.image_token_id, lo
    cfg["vocab_size"] = cfg["text_config"]["vocab_size"] = (lo + nb["total"] + 15) // 8 * 8
    rng = np.random.default_rng(4)
    actions = rng.uniform(-1, 1, size=(2, 7))                                  # one sample, chunk of 2 actions
    img = (rng.random((224, 224, 3)) * 255).astype(np.uint8)
    out = p(images=[img], text=["pick up the cup"], unnorm_key="bridge", suffix_actions=actions)
    ids, tt, labels = out["input_ids"], out["token_type_ids"], out["labels"]
    n_suffix = 2 * 3 + 1                                                        # 3 tokens per action + EOS
    assert int(tt.sum()) == n_suffix and (tt[0, -n_suffix:] == 1).all() and (tt[0, :-n_suffix] == 0).all()
    assert (labels[0, :-n_suffix] == -100).all() and torch.equal(labels[0, -n_suffix:], ids[0, -n_suffix:])
    assert int(ids[0, -1]) == 1                                                 # EOS
    want = T.encode(actions, tk.bin_policy, nb) + lo                            # the suffix ids ARE the grid ids of the actions
    assert np.array_equal(ids[0, -n_suffix:-1].numpy().reshape(2, 3), want)
    sd = synth_state_dict(cfg, seed=0)
    m = SpatialVLAForConditionalGeneration(cfg, sd, ops=RefOps())
    res = m(input_ids=ids, pixel_values=out["pixel_values"], intrinsic=out["intrinsic"], attention_mask=out["attention_mask"],
            token_type_ids=tt, labels=labels)
    assert res.label_rows.numel() == n_suffix and torch.isfinite(res.loss)
    ref_loss, rows, lab, _ = R.forward_loss_ref(sd, cfg, ids, out["pixel_values"], out["intrinsic"], labels, token_type_ids=tt,
                                                attention_mask=out["attention_mask"], force_head=m.engine.last_router_head)
    assert torch.equal(res.label_rows, rows) and abs(float(res.loss) - float(ref_loss)) < 1e-2
    met = m.action_metrics(res, actions, tk)
    assert 0.0 <= met["accuracy"] <= 1.0 and met["l1_loss"] >= 0.0
    # a perfect prediction gives accuracy 1 and an L1 distance of at most one bin width
    res.row_argmax = res.row_labels.clone()
    met = m.action_metrics(res, actions, tk)
    assert met["accuracy"] == 1.0 and met["translation_accuracy"] == 1.0 and met["gripper_accuracy"] == 1.0 and met["l1_loss"] < 0.3


def test_spatial_embedding_adaption_matches_reference_golden():
    """Fine-tune-time re-gridding (model/action_tokenizer.py:372-430): golden minted from the live reference
    (oracle/gen_golden.py adaption) -- new bin edges exactly, re-sampled embeddings to 1e-6, NaN pattern (outside the old hull) equal."""
    import os
    from oracle.gen_golden import ADAPT_BINS, ADAPT_GS0, ADAPT_GS1
    from spatialvla_b200 import SpatialActionTokenizer
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "embedding_adaption.npz"))
    tk = SpatialActionTokenizer(FakeTokenizer(1000), ADAPT_BINS, gs_params=ADAPT_GS0, min_sigma=0.1)
    emb = torch.nn.Embedding(g["before"].shape[0], g["before"].shape[1])
    emb.weight.data.copy_(torch.from_numpy(g["before"]))
    edges_before = tk._edges.copy()
    tk.spatial_embedding_adaption(ADAPT_GS1, emb, min_sigma=0.2, adpt_feature=True)
    for bt in ("translation", "rotation"):
        for k in ADAPT_BINS[bt]:
            assert np.array_equal(np.asarray(tk.bin_policy[bt][k]), g[f"edge_{k}"]), k
    assert not np.array_equal(tk._edges, edges_before)                 # the device-side edge table follows the new policy
    a, b = emb.weight.data.numpy(), g["after"]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.allclose(np.nan_to_num(a), np.nan_to_num(b), atol=1e-6)
    # adpt_feature=False only swaps the grids
    tk2 = SpatialActionTokenizer(FakeTokenizer(1000), ADAPT_BINS, gs_params=ADAPT_GS0, min_sigma=0.1)
    emb2 = torch.nn.Embedding(g["before"].shape[0], g["before"].shape[1])
    emb2.weight.data.copy_(torch.from_numpy(g["before"]))
    tk2.spatial_embedding_adaption(ADAPT_GS1, emb2, min_sigma=0.2)
    assert np.array_equal(emb2.weight.data.numpy(), g["before"]) and np.array_equal(tk2._edges, tk._edges)


def _adaption_with_ops(ops):
    import os
    from oracle.gen_golden import ADAPT_BINS, ADAPT_GS0, ADAPT_GS1
    from spatialvla_b200 import SpatialActionTokenizer
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "embedding_adaption.npz"))
    tk = SpatialActionTokenizer(FakeTokenizer(1000), ADAPT_BINS, gs_params=ADAPT_GS0, min_sigma=0.1)
    emb = torch.nn.Embedding(g["before"].shape[0], g["before"].shape[1])
    emb.weight.data.copy_(torch.from_numpy(g["before"]))
    tk.spatial_embedding_adaption(ADAPT_GS1, emb, min_sigma=0.2, adpt_feature=True, ops=ops)
    a, b = emb.weight.data.numpy(), g["after"]
    return a, b


def test_spatial_embedding_adaption_plan_and_gather_match_reference_golden():
    """The device formulation (host: Qhull triangulation + point location = `adaption_plan`; device: 4-row barycentric gather)
    run through the torch op re-statement against the golden the LIVE reference produced with scipy griddata: same NaN pattern,
    values to 1e-6 (the reference's interpolation is Delaunay-linear, which the plan reproduces, not trilinear)."""
    from oracle.ops_ref import RefOps
    a, b = _adaption_with_ops(RefOps())
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.isnan(b).sum() < b.size
    assert np.allclose(np.nan_to_num(a), np.nan_to_num(b), atol=1e-6)


@pytest.mark.gpu
def test_spatial_embedding_adaption_on_gpu_matches_reference_golden(cuda_device):
    from spatialvla_b200.ops import CudaOps
    ops = CudaOps(cuda_device)
    n0 = ops.launch_count()
    a, b = _adaption_with_ops(ops)
    assert ops.launch_count() - n0 == 2                    # one gather per block (translation, rotation)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    assert np.allclose(np.nan_to_num(a), np.nan_to_num(b), atol=1e-6)


def test_prompt_id_cache_per_instruction():
    """Serving loop: the same instruction every control step -> the tokenizer runs once per distinct prompt;
    results are identical and independent copies; training samples (suffix) bypass the cache."""
    p = make_processor()
    calls = []
    real = p.tokenizer.__class__.__call__

    class Counting(p.tokenizer.__class__):
        def __call__(self, *a, **k):
            calls.append(a[0])
            return real(self, *a, **k)
    p.tokenizer.__class__ = Counting
    img = (np.random.default_rng(0).random((224, 224, 3)) * 255).astype(np.uint8)
    a = p(images=[img], text=["pick up the cup"], unnorm_key="bridge")
    b = p(images=[img], text=["pick up the cup"], unnorm_key="bridge")
    c = p(images=[img], text=["open drawer"], unnorm_key="bridge")
    assert len(calls) == 2 and torch.equal(a["input_ids"], b["input_ids"]) and not torch.equal(a["input_ids"][0, 257:260], c["input_ids"][0, 257:260])
    b["input_ids"][0, 0] = -7                                  # a returned tensor is a copy: the cache entry is not corrupted
    d = p(images=[img], text=["pick up the cup"], unnorm_key="bridge")
    assert len(calls) == 2 and int(d["input_ids"][0, 0]) == int(a["input_ids"][0, 0])
    p.prompt_cache_size = 1                                    # LRU bound
    p(images=[img], text=["close drawer"], unnorm_key="bridge")
    assert len(p._prompt_cache) == 1 and len(calls) == 3
    p(images=[img], text=["pick up the cup"], unnorm_key="bridge")
    assert len(calls) == 4
    p.prompt_cache_size = 0                                    # disabled
    p(images=[img], text=["pick up the cup"], unnorm_key="bridge")
    assert len(calls) == 5
