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
