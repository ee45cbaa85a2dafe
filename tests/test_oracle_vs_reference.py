"""CPU, build container only (skipped where /root/reference is absent): pins the restatement and the state_dict
layout against the LIVE reference through the compat shim."""
import pytest
import torch

from oracle import compat

pytestmark = pytest.mark.skipif(not compat.reference_available(), reason="reference tree not present")


def test_state_dict_spec_matches_reference():
    from spatialvla_b200.configs import get_config_dict
    from spatialvla_b200.weights import state_dict_spec
    cfg = get_config_dict("tiny")
    ref_sd = compat.build_reference_model(cfg).state_dict()
    spec = state_dict_spec(cfg)
    assert set(spec) == set(ref_sd)
    for k, shp in spec.items():
        assert tuple(ref_sd[k].shape) == tuple(shp), k


def test_tokenizer_live_reference_random_grids():
    import numpy as np
    from oracle import tokenizer_ref as T
    from oracle.gen_golden import FakeHFTokenizer
    _, tok_mod, _, _ = compat.import_reference()
    nb = {"translation": {"theta_bins": 5, "phi_bins": 7, "r_bins": 3}, "rotation": {"roll_bins": 4, "pitch_bins": 6, "yaw_bins": 2},
          "gripper": 2}
    gs = {k: {"mu": 0.1 * i - 0.3, "sigma": 0.2 + 0.1 * i} for i, k in enumerate(("theta", "phi", "r", "roll", "pitch", "yaw"))}
    tk = tok_mod.SpatialActionTokenizer(FakeHFTokenizer(1000), num_bins=nb, gs_params=gs, min_sigma=0.3)
    acts = np.random.default_rng(1).uniform(-1.2, 1.2, size=(5000, 7))
    ref = np.vectorize(lambda s: int(s[7:12]))(tk(acts))
    assert np.array_equal(T.encode(acts, tk.bin_policy, nb), ref)
    ids = np.stack([np.arange(105) % 105, 105 + np.arange(105) % 48, 153 + np.arange(105) % 2], 1)
    assert np.array_equal(T.decode(ids, tk.bin_policy, nb), tk.decode_token_ids_to_actions(ids + 1000))


def test_spatial_embedding_adaption_matches_live_reference():
    """Fine-tune-time re-gridding (model/action_tokenizer.py:372-430): new bin policy + scattered-linear
    re-sampling of the spatial embeddings, against the live reference on the same grids / embeddings."""
    import numpy as np
    from oracle.gen_golden import FakeHFTokenizer
    from spatialvla_b200.action_tokenizer import SpatialActionTokenizer
    _, tok_mod, _, _ = compat.import_reference()
    nb = {"translation": {"theta_bins": 4, "phi_bins": 5, "r_bins": 3}, "rotation": {"roll_bins": 3, "pitch_bins": 4, "yaw_bins": 2},
          "gripper": 2}
    gs0 = {k: {"mu": 0.05 * i - 0.1, "sigma": 0.3 + 0.05 * i} for i, k in enumerate(("theta", "phi", "r", "roll", "pitch", "yaw"))}
    gs1 = {k: {"mu": -0.04 * i + 0.1, "sigma": 0.45 - 0.03 * i} for i, k in enumerate(("theta", "phi", "r", "roll", "pitch", "yaw"))}
    n_tok = 4 * 5 * 3 + 3 * 4 * 2 + 2
    w = torch.randn(n_tok, 6, generator=torch.Generator().manual_seed(0))
    ref_tk = tok_mod.SpatialActionTokenizer(FakeHFTokenizer(1000), num_bins=nb, gs_params=gs0, min_sigma=0.1)
    our_tk = SpatialActionTokenizer(FakeHFTokenizer(1000), nb, gs_params=gs0, min_sigma=0.1)
    ref_emb, our_emb = torch.nn.Embedding(n_tok, 6), torch.nn.Embedding(n_tok, 6)
    ref_emb.weight.data.copy_(w)
    our_emb.weight.data.copy_(w)
    ref_tk.spatial_embedding_adaption(gs1, ref_emb, min_sigma=0.2, adpt_feature=True)
    our_tk.spatial_embedding_adaption(gs1, our_emb, min_sigma=0.2, adpt_feature=True)
    for bt in ("translation", "rotation"):
        for k in nb[bt]:
            assert np.array_equal(np.asarray(our_tk.bin_policy[bt][k]), np.asarray(ref_tk.bin_policy[bt][k])), k
    a, b = our_emb.weight.data, ref_emb.weight.data
    assert torch.equal(torch.isnan(a), torch.isnan(b))                 # points outside the old hull are NaN in both (scipy griddata)
    assert torch.allclose(torch.nan_to_num(a), torch.nan_to_num(b), atol=1e-6)
    assert not torch.equal(torch.nan_to_num(a[:60]), torch.nan_to_num(w[:60]))      # the re-sampling did something
    assert torch.equal(a[-2:], w[-2:])                                              # gripper rows untouched
