"""spatialvla_b200 -- B200-native (sm_100a) implementation of SpatialVLA's action-prediction path.

Public surface mirrors the reference's `model` package (model/__init__.py:16-30): SpatialVLAConfig,
SpatialVLAProcessor, SpatialActionTokenizer, SpatialVLAForConditionalGeneration.  Importing the package is cheap and
GPU-free; the CUDA library is loaded on first use of a compute object and there is no CPU fallback."""
from .configs import CANONICAL_4B_224, TINY, get_config_dict, default_intrinsic_224  # noqa: F401

__all__ = ["SpatialVLAConfig", "SpatialVLAProcessor", "SpatialActionTokenizer", "SpatialVLAForConditionalGeneration",
           "get_config_dict"]


def __getattr__(name):
    if name == "SpatialVLAConfig":
        from .configuration_spatialvla import SpatialVLAConfig
        return SpatialVLAConfig
    if name == "SpatialVLAProcessor":
        from .processing_spatialvla import SpatialVLAProcessor
        return SpatialVLAProcessor
    if name == "SpatialActionTokenizer":
        from .action_tokenizer import SpatialActionTokenizer
        return SpatialActionTokenizer
    if name == "SpatialVLAForConditionalGeneration":
        from .modeling_spatialvla import SpatialVLAForConditionalGeneration
        return SpatialVLAForConditionalGeneration
    raise AttributeError(name)
