"""`SpatialVLAConfig` -- same fields, defaults and sub-config handling as the reference
(model/configuration_spatialvla.py:22-102), so a config.json written by the reference loads here unchanged.
`to_engine_dict()` flattens it into the plain nested dict the engine consumes."""
from __future__ import annotations

import warnings

from transformers import CONFIG_MAPPING, AutoConfig
from transformers.configuration_utils import PretrainedConfig


class SpatialVLAConfig(PretrainedConfig):
    model_type = "spatialvla"
    sub_configs = {"text_config": AutoConfig, "vision_config": AutoConfig, "vision_zoe_config": AutoConfig}

    def __init__(self, vision_config=None, text_config=None, ignore_index=-100, image_token_index=256000,
                 vocab_size=257152, projection_dim=2048, hidden_size=2048, vision_zoe_config=None,
                 action_token_begin_idx=None, spatial_token_num=259, use_spatial_token=False, ego3d_patch_reso=4,
                 n_freqs=8, use_vision_zoe=True, **kwargs):
        self._ignore_index = ignore_index
        self.image_token_index = image_token_index
        self._vocab_size = vocab_size
        self.projection_dim = projection_dim
        self.hidden_size = hidden_size
        self.is_encoder_decoder = False

        if isinstance(vision_config, dict):
            vision_config = dict(vision_config)
            vision_config.setdefault("model_type", "siglip_vision_model")
            vision_config = CONFIG_MAPPING[vision_config["model_type"]](**vision_config)
        elif vision_config is None:
            vision_config = CONFIG_MAPPING["siglip_vision_model"](
                intermediate_size=4096, hidden_size=1152, patch_size=14, image_size=224, num_hidden_layers=27,
                num_attention_heads=16, vocab_size=257152, vision_use_head=False)
        self.vision_config = vision_config

        if isinstance(text_config, dict):
            text_config = dict(text_config)
            text_config.setdefault("model_type", "gemma2")
            text_config = CONFIG_MAPPING[text_config["model_type"]](**text_config)
        elif text_config is None:
            text_config = CONFIG_MAPPING["gemma2"](hidden_size=2048, num_hidden_layers=18, intermediate_size=16384,
                                                   num_attention_heads=8, num_key_value_heads=1,
                                                   is_encoder_decoder=False, vocab_size=vocab_size)
        self.text_config = text_config
        self.text_config.num_image_tokens = (self.vision_config.image_size // self.vision_config.patch_size) ** 2
        self.vision_config.projection_dim = projection_dim

        if isinstance(vision_zoe_config, dict):
            vision_zoe_config = dict(vision_zoe_config)
            vision_zoe_config.setdefault("model_type", "zoedepth")
            vision_zoe_config = CONFIG_MAPPING[vision_zoe_config["model_type"]](**vision_zoe_config)
        self.vision_zoe_config = vision_zoe_config

        self.action_token_begin_idx = action_token_begin_idx
        self.spatial_token_num = spatial_token_num
        self.use_spatial_token = use_spatial_token
        self.ego3d_patch_reso = ego3d_patch_reso
        self.n_freqs = n_freqs
        self.use_vision_zoe = use_vision_zoe
        super().__init__(**kwargs)

    @property
    def ignore_index(self):
        warnings.warn("The `ignore_index` attribute is deprecated and will be removed in v4.47.", FutureWarning)
        return self._ignore_index

    @ignore_index.setter
    def ignore_index(self, value):
        self._ignore_index = value

    def to_dict(self):
        output = super().to_dict()
        output.pop("_ignore_index", None)
        return output

    # ---- bridge to the engine
    def to_engine_dict(self) -> dict:
        t = self.text_config.to_dict()
        if "rope_theta" not in t or t["rope_theta"] is None:
            rp = t.get("rope_parameters") or {}
            t["rope_theta"] = float(rp.get("rope_theta", getattr(self.text_config, "rope_theta", 10000.0) or 10000.0))
        d = {
            "vision_config": self.vision_config.to_dict(), "text_config": t,
            "image_token_index": self.image_token_index, "action_token_begin_idx": self.action_token_begin_idx,
            "spatial_token_num": self.spatial_token_num, "use_spatial_token": self.use_spatial_token,
            "ego3d_patch_reso": self.ego3d_patch_reso, "n_freqs": self.n_freqs, "use_vision_zoe": self.use_vision_zoe,
            # read by the labelled forward: labels on pad-token inputs are masked (model/modeling_spatialvla.py:389-397, where
            # pad_token_id None becomes -1 at :192)
            "pad_token_id": getattr(self, "pad_token_id", None), "ignore_index": getattr(self, "_ignore_index", -100),
            "eos_token_id": getattr(self, "eos_token_id", None),
        }
        if self.use_vision_zoe and self.vision_zoe_config is not None:
            z = self.vision_zoe_config.to_dict()
            bb = self.vision_zoe_config.backbone_config
            z["backbone_config"] = bb.to_dict() if hasattr(bb, "to_dict") else dict(bb)
            d["vision_zoe_config"] = z
        return d
