"""Plain-dict model configurations (kwargs of `SpatialVLAConfig`).

`CANONICAL_4B_224` is the contract of SURVEY.md Appendix A (SpatialVLA-4B-224: SigLIP-So400m/14 + ZoeDepth-NK
(BEiT-L/16 @384) + Gemma2-2B, vocab 265 347 with an 8 194-token spatial action slice).  `TINY` keeps every
*kernel-visible* constant (head dims 72 / 64 / 32 / 256, 224x224 image -> 256 patches, 384x384 ZoeDepth input,
two metric-bin heads with the batch-level router, 8 194 action tokens) but shrinks widths and depths so that the
fp32 reference finishes in seconds on CPU; it is the configuration the golden vectors in tests/golden were
minted on (oracle/gen_golden.py).
"""
from __future__ import annotations

import copy

_BINS_NK = [
    {"name": "nyu", "n_bins": 64, "min_depth": 1e-3, "max_depth": 10.0},
    {"name": "kitti", "n_bins": 64, "min_depth": 1e-3, "max_depth": 80.0},
]

CANONICAL_4B_224 = {
    "vision_config": {
        "model_type": "siglip_vision_model", "hidden_size": 1152, "intermediate_size": 4304,
        "num_hidden_layers": 27, "num_attention_heads": 16, "patch_size": 14, "image_size": 224,
        "vision_use_head": False, "layer_norm_eps": 1e-6, "hidden_act": "gelu_pytorch_tanh",
        "projection_dim": 2304,
    },
    "text_config": {
        "model_type": "gemma2", "hidden_size": 2304, "intermediate_size": 9216, "num_hidden_layers": 26,
        "num_attention_heads": 8, "num_key_value_heads": 4, "head_dim": 256, "vocab_size": 265347,
        "sliding_window": 4096, "query_pre_attn_scalar": 256, "attn_logit_softcapping": 50.0,
        "final_logit_softcapping": 30.0, "rms_norm_eps": 1e-6, "hidden_activation": "gelu_pytorch_tanh",
        "attention_bias": False, "tie_word_embeddings": False, "max_position_embeddings": 8192,
    },
    "vision_zoe_config": {
        "model_type": "zoedepth",
        "backbone_config": {
            "model_type": "beit", "image_size": 384, "patch_size": 16, "hidden_size": 1024,
            "num_hidden_layers": 24, "intermediate_size": 4096, "num_attention_heads": 16,
            "use_relative_position_bias": True, "reshape_hidden_states": False,
            "out_features": ["stage6", "stage12", "stage18", "stage24"],
        },
        "neck_hidden_sizes": [256, 512, 1024, 1024], "fusion_hidden_size": 256, "bottleneck_features": 256,
        "num_relative_features": 32, "bin_embedding_dim": 128, "num_attractors": [16, 8, 4, 1],
        "bin_centers_type": "softplus", "readout_type": "project", "reassemble_factors": [4, 2, 1, 0.5],
        "bin_configurations": _BINS_NK, "num_patch_transformer_layers": 4,
        "patch_transformer_hidden_size": 128, "patch_transformer_intermediate_size": 1024,
        "patch_transformer_num_attention_heads": 4, "min_temp": 0.0212, "max_temp": 50.0,
    },
    "image_token_index": 257152, "vocab_size": 265347, "projection_dim": 2304, "hidden_size": 2304,
    "action_token_begin_idx": 257153, "spatial_token_num": 8194, "use_spatial_token": True,
    "ego3d_patch_reso": 2, "n_freqs": 8, "use_vision_zoe": True,
    "pad_token_id": 0, "bos_token_id": 2, "eos_token_id": 1,
}

TINY = {
    "vision_config": {
        "model_type": "siglip_vision_model", "hidden_size": 144, "intermediate_size": 272,
        "num_hidden_layers": 2, "num_attention_heads": 2, "patch_size": 14, "image_size": 224,
        "vision_use_head": False, "layer_norm_eps": 1e-6, "hidden_act": "gelu_pytorch_tanh",
        "projection_dim": 512,
    },
    "text_config": {
        "model_type": "gemma2", "hidden_size": 512, "intermediate_size": 1024, "num_hidden_layers": 3,
        "num_attention_heads": 4, "num_key_value_heads": 2, "head_dim": 256, "vocab_size": 9216,
        "sliding_window": 4096, "query_pre_attn_scalar": 256, "attn_logit_softcapping": 50.0,
        "final_logit_softcapping": 30.0, "rms_norm_eps": 1e-6, "hidden_activation": "gelu_pytorch_tanh",
        "attention_bias": False, "tie_word_embeddings": False, "max_position_embeddings": 8192,
    },
    "vision_zoe_config": {
        "model_type": "zoedepth",
        "backbone_config": {
            "model_type": "beit", "image_size": 384, "patch_size": 16, "hidden_size": 128,
            "num_hidden_layers": 4, "intermediate_size": 256, "num_attention_heads": 2,
            "use_relative_position_bias": True, "reshape_hidden_states": False,
            "out_features": ["stage1", "stage2", "stage3", "stage4"],
        },
        "neck_hidden_sizes": [64, 64, 128, 128], "fusion_hidden_size": 64, "bottleneck_features": 64,
        "num_relative_features": 32, "bin_embedding_dim": 128, "num_attractors": [16, 8, 4, 1],
        "bin_centers_type": "softplus", "readout_type": "project", "reassemble_factors": [4, 2, 1, 0.5],
        "bin_configurations": _BINS_NK, "num_patch_transformer_layers": 4,
        "patch_transformer_hidden_size": 128, "patch_transformer_intermediate_size": 256,
        "patch_transformer_num_attention_heads": 4, "min_temp": 0.0212, "max_temp": 50.0,
    },
    "image_token_index": 1021, "vocab_size": 9216, "projection_dim": 512, "hidden_size": 512,
    "action_token_begin_idx": 1022, "spatial_token_num": 8194, "use_spatial_token": True,
    "ego3d_patch_reso": 2, "n_freqs": 8, "use_vision_zoe": True,
    "pad_token_id": 0, "bos_token_id": 2, "eos_token_id": 1,
}


def get_config_dict(name: str) -> dict:
    table = {"4b-224": CANONICAL_4B_224, "canonical": CANONICAL_4B_224, "tiny": TINY}
    return copy.deepcopy(table[name])


# Default camera intrinsics of the reference (scripts/intrinsics.json "default": fx=fy=623.588, cx=319.501,
# cy=239.545 at 640x480) rescaled to the 224x224 model input as model/processing_spatialvla.py:91-95 does.
def default_intrinsic_224():
    fx, fy, cx, cy, W, H = 623.588, 623.588, 319.501, 239.545, 640.0, 480.0
    sx, sy = 224.0 / W, 224.0 / H
    return [[fx * sx, 0.0, cx * sx], [0.0, fy * sy, cy * sy], [0.0, 0.0, 1.0]]
