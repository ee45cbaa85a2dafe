"""`SpatialVLAProcessor` with the reference's call surface (model/processing_spatialvla.py:41-254).

Host-side work (prompt strings, PIL resize through the SigLIP image processor, text tokenisation, per-dataset
intrinsics) stays in Python exactly as in the reference; the action grid arithmetic goes through the CUDA
tokenizer.  It duck-types its collaborators (any HF-style `image_processor` / `tokenizer` objects work), so it
needs no hub access.  Additions for batched serving: `decode_actions_batch` (all rows, the reference decodes row 0
only -- kept as `decode_actions`)."""
from __future__ import annotations

import logging
from typing import Dict, List, Optional

import numpy as np
import torch

from .action_tokenizer import SpatialActionTokenizer

logger = logging.getLogger(__name__)

IMAGE_TOKEN = "<image>"
EXTRA_TOKENS = [f"<loc{i:0>4}>" for i in range(1024)] + [f"<seg{i:0>3}>" for i in range(128)]


# Keyword routing of HF `ProcessorMixin._merge_kwargs` for PaliGemmaProcessorKwargs (model/processing_spatialvla.py:113-117):
# `images_kwargs` go to the image processor, `text_kwargs` to the tokenizer, `return_tensors` (common) to both.  The reference's
# training call (data/dataset.py:134-142) passes do_normalize=False / max_length / truncation / padding in ONE kwargs bag.
IMAGES_KWARGS = frozenset({
    "do_resize", "size", "size_divisor", "crop_size", "resample", "do_rescale", "rescale_factor", "do_normalize", "image_mean",
    "image_std", "do_pad", "pad_size", "do_center_crop", "data_format", "input_data_format", "do_convert_rgb", "device"})
TEXT_KWARGS = frozenset({
    "text_pair", "text_target", "text_pair_target", "add_special_tokens", "padding", "truncation", "max_length", "stride",
    "is_split_into_words", "pad_to_multiple_of", "return_token_type_ids", "return_attention_mask", "return_overflowing_tokens",
    "return_special_tokens_mask", "return_offsets_mapping", "return_length", "verbose", "padding_side", "return_mm_token_type_ids",
    "suffix"})


def split_processor_kwargs(kwargs):
    """-> (images_kwargs, text_kwargs); keys HF does not know are dropped with a warning, as `_merge_kwargs` does."""
    im, tx = {}, {}
    for k, v in kwargs.items():
        if k in IMAGES_KWARGS:
            im[k] = v
        elif k in TEXT_KWARGS:
            tx[k] = v
        else:
            logger.warning("SpatialVLAProcessor: keyword argument %r is not a text / image processing option; ignored", k)
    return im, tx


def build_string_from_input(prompt, bos_token, image_seq_len, image_token, num_images):
    """HF paligemma/processing_paligemma.py:76-95 -- '<image>'*n + bos + prompt + newline"""
    return f"{image_token * image_seq_len * num_images}{bos_token}{prompt}\n"


class BatchFeature(dict):
    """Minimal dict-like batch with `.to()` (dtype applies to floating tensors only, like HF's BatchFeature)."""

    def to(self, *args, **kwargs):
        out = BatchFeature()
        for k, v in self.items():
            if isinstance(v, torch.Tensor):
                if v.is_floating_point():
                    out[k] = v.to(*args, **kwargs)
                else:
                    dev = kwargs.get("device")
                    for a in args:
                        if isinstance(a, (str, torch.device)):
                            dev = a
                    out[k] = v.to(device=dev) if dev is not None else v
            else:
                out[k] = v
        return out

    def __getattr__(self, item):
        try:
            return self[item]
        except KeyError as e:
            raise AttributeError(item) from e


class SpatialVLAProcessor:
    attributes = ["image_processor", "tokenizer"]

    def __init__(self, image_processor=None, tokenizer=None, chat_template=None, statistics: Optional[dict] = None,
                 bin_policy=None, intrinsic_config=None, action_config=None, num_obs_steps=1, obs_delta=1,
                 action_chunk_size=1, min_sigma=0.0, **kwargs):
        if image_processor is None:
            raise ValueError("You need to specify an `image_processor`.")
        if tokenizer is None:
            raise ValueError("You need to specify a `tokenizer`.")
        if not hasattr(image_processor, "image_seq_length"):
            raise ValueError("Image processor is missing an `image_seq_length` attribute.")
        self.image_processor, self.tokenizer, self.chat_template = image_processor, tokenizer, chat_template
        self.image_seq_length = image_processor.image_seq_length
        if not hasattr(tokenizer, "image_token"):
            tokenizer.add_special_tokens({"additional_special_tokens": [IMAGE_TOKEN]})
            self.image_token_id = tokenizer.convert_tokens_to_ids(IMAGE_TOKEN)
        else:
            self.image_token_id = tokenizer.image_token_id
        tokenizer.add_tokens(EXTRA_TOKENS)
        tokenizer.add_bos_token = False
        tokenizer.add_eos_token = False

        # prompt-id cache: a control loop sends the same instruction every step, so the tokenizer output of
        # an inference prompt is kept per (prompt strings, tokenizer kwargs); training samples (suffix) are never cached
        self.prompt_cache_size = 256
        self._prompt_cache = {}
        self.statistics = statistics if statistics else {}
        self.bin_policy = bin_policy
        self.min_sigma = min_sigma
        self.intrinsic_config = intrinsic_config
        self.action_config = action_config
        self.num_obs_steps = num_obs_steps
        self.obs_delta = obs_delta
        self.action_chunk_size = action_chunk_size
        self.dataset_intrinsics = {}
        height, width = image_processor.size["height"], image_processor.size["width"]
        for k, v in intrinsic_config.items():          # model/processing_spatialvla.py:91-95
            K = torch.tensor(v["intrinsic"]).float()
            K[:2] *= torch.tensor([width / v["width"], height / v["height"]])[:, None]
            self.dataset_intrinsics[k] = K
        self.action_tokenizer = SpatialActionTokenizer(
            tokenizer=tokenizer, num_bins=action_config["num_bins"], bin_policy=bin_policy,
            use_spherical=action_config["use_spherical"], min_sigma=min_sigma)

    def __call__(self, images=None, text=None, unnorm_key: Optional[str] = None, suffix_actions=None,
                 return_tensors="pt", suffix=None, **kwargs) -> BatchFeature:
        images_kwargs, text_kwargs = split_processor_kwargs(kwargs)
        if suffix_actions is not None:
            action_tokens = self.action_tokenizer(suffix_actions)
            suffix = "".join(action_tokens.flatten())
            text_kwargs.pop("suffix", None)
        elif suffix is None:
            suffix = text_kwargs.pop("suffix", None)
        return_token_type_ids = suffix is not None
        if images is None:
            raise ValueError("`images` are expected as arguments to a `PaliGemmaProcessor` instance.")
        if text is None:
            text = ""
        if isinstance(text, str):
            text = [text]
        if not isinstance(images, (list, tuple)):
            images = [[images]]
        elif len(images) and not isinstance(images[0], (list, tuple)):
            images = [[im] for im in images]
        if not any(IMAGE_TOKEN in s for s in text):
            if len(images) != len(text):
                raise ValueError(f"Received {len(images)} images for {len(text)} prompts. Each prompt should be "
                                 "associated with an image or list of images.")
            if suffix is not None and isinstance(suffix, str):
                suffix = [suffix]
            if suffix is not None:
                suffix = [sfx + self.tokenizer.eos_token for sfx in suffix]
            input_strings = [build_string_from_input(p, self.tokenizer.bos_token, self.image_seq_length, IMAGE_TOKEN,
                                                     len(il)) for p, il in zip(text, images)]
        else:
            input_strings = []
            for sample in text:
                ex = sample.replace(IMAGE_TOKEN, IMAGE_TOKEN * self.image_seq_length)
                idx = ex.rfind(IMAGE_TOKEN)
                idx = idx + len(IMAGE_TOKEN) if idx != -1 else 0
                input_strings.append(ex[:idx] + self.tokenizer.bos_token + ex[idx:] + "\n")
        flat = [im for il in images for im in il]
        pixel_values = self._device_pixel_values(flat, images_kwargs)
        if pixel_values is None:
            pixel_values = self.image_processor(flat, return_tensors=return_tensors, **images_kwargs)["pixel_values"]
        if text_kwargs.get("max_length") is not None:          # the limit is quoted without the image tokens (:176-178)
            text_kwargs["max_length"] = text_kwargs["max_length"] + self.image_seq_length
        inputs = self._tokenize(input_strings, suffix, return_token_type_ids, return_tensors, text_kwargs)
        intrinsic = self.dataset_intrinsics[unnorm_key] if unnorm_key in self.dataset_intrinsics \
            else self.dataset_intrinsics["default"]
        data = {**inputs, "pixel_values": pixel_values, "intrinsic": intrinsic}
        if return_token_type_ids:
            data["labels"] = inputs["input_ids"].masked_fill(inputs["token_type_ids"] == 0, -100)
        return BatchFeature(data)

    # ---- device-side image path
    def enable_device_images(self, ops=None, device="cuda:0"):
        """Resize / rescale / (normalise) the uint8 camera frames on the GPU (csrc/image_ops.cu: Pillow's bicubic resampler bit for
        bit) instead of in PIL on the host: `__call__` then returns `pixel_values` already resident on the device.  Frames that are
        not uint8 HWC arrays of one common size keep the host image processor."""
        if ops is None:
            from .ops import CudaOps
            ops = CudaOps(device)
        self._image_ops, self._device_processors = ops, {}
        return self

    def _device_pixel_values(self, flat, images_kwargs):
        ops = getattr(self, "_image_ops", None)
        if ops is None or not flat:
            return None
        arrs = []
        for im in flat:
            a = im if isinstance(im, torch.Tensor) else np.asarray(im)
            if a.dtype not in (np.uint8, torch.uint8) or a.ndim != 3 or a.shape[-1] != 3 or tuple(a.shape) != tuple(flat[0].shape if hasattr(flat[0], "shape") else np.asarray(flat[0]).shape):
                return None
            arrs.append(a)
        ip = self.image_processor
        opt = lambda k, d: images_kwargs.get(k, getattr(ip, k, d))      # noqa: E731
        if not opt("do_resize", True) or images_kwargs.get("size") is not None or images_kwargs.get("resample") is not None:
            return None
        key = (bool(opt("do_rescale", True)), float(opt("rescale_factor", 1 / 255)), bool(opt("do_normalize", False)),
               tuple(opt("image_mean", (0.5, 0.5, 0.5)) or (0.5, 0.5, 0.5)), tuple(opt("image_std", (0.5, 0.5, 0.5)) or (0.5, 0.5, 0.5)))
        if key not in self._device_processors:
            from .image_processing import DeviceImageProcessor
            self._device_processors[key] = DeviceImageProcessor(ops, (ip.size["height"], ip.size["width"]), rescale_factor=key[1],
                                                               do_rescale=key[0], do_normalize=key[2], mean=key[3], std=key[4])
        batch = torch.stack(arrs) if isinstance(arrs[0], torch.Tensor) else np.stack(arrs)
        return self._device_processors[key](batch)

    def _tokenize(self, input_strings, suffix, return_token_type_ids, return_tensors, kwargs):
        key = None
        if suffix is None and self.prompt_cache_size > 0:
            try:
                key = (tuple(input_strings), return_tensors, tuple(sorted(kwargs.items())))
                hash(key)
            except TypeError:
                key = None
        if key is not None and key in self._prompt_cache:
            hit = self._prompt_cache.pop(key)
            self._prompt_cache[key] = hit                      # most recently used last
            return {k: (v.clone() if hasattr(v, "clone") else v) for k, v in hit.items()}
        inputs = self.tokenizer(input_strings, text_pair=suffix, return_token_type_ids=return_token_type_ids,
                                return_tensors=return_tensors, **kwargs)
        if key is not None:
            self._prompt_cache[key] = {k: (v.clone() if hasattr(v, "clone") else v) for k, v in dict(inputs).items()}
            while len(self._prompt_cache) > self.prompt_cache_size:
                self._prompt_cache.pop(next(iter(self._prompt_cache)))
        return dict(inputs)

    def batch_decode(self, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.tokenizer.decode(*args, **kwargs)

    @property
    def model_input_names(self):
        return list(dict.fromkeys(list(self.tokenizer.model_input_names) + list(self.image_processor.model_input_names)))

    # ---- action decoding
    def _unnormalize(self, normalized, unnorm_key):
        if unnorm_key is None:
            logger.warning("unnorm_key None is not in statistics, use next one")
            unnorm_key = next(iter(self.statistics.keys()))
        st = self.statistics[unnorm_key]["action"]
        dim = len(st["q01"])
        mask = np.array(st.get("mask", np.ones(dim)), dtype=bool)
        hi, lo = np.array(st["q99"]), np.array(st["q01"])
        return np.where(mask, 0.5 * (normalized + 1) * (hi - lo) + lo, normalized)

    def decode_actions(self, generation_outputs: torch.Tensor, unnorm_key: Optional[str] = None) -> Dict[str, np.ndarray]:
        """Row 0 only, exactly like the reference (model/processing_spatialvla.py:216-254)."""
        n_tok = 3
        ids = generation_outputs[0, : n_tok * self.action_chunk_size].detach().cpu().long().numpy()
        if ids.shape[0] < n_tok * self.action_chunk_size:
            logger.warning("Padding zero action!")
            ids = np.concatenate([ids, np.zeros(n_tok * self.action_chunk_size - ids.shape[0], dtype=np.longlong)])
        ids = ids.reshape(-1, n_tok)
        normalized = self.action_tokenizer.decode_token_ids_to_actions(ids)
        return {"actions": self._unnormalize(normalized, unnorm_key), "action_ids": ids}

    def decode_actions_device(self, generation_outputs: torch.Tensor, unnorm_key: Optional[str] = None):
        """All rows, device-resident: CUDA ids (B, >= 3*chunk) -> {'actions': float64 CUDA tensor
        (B, chunk, 7), 'action_ids': int64 (B, chunk, 3)} without a host round trip.  Same arithmetic as `decode_actions`:
        the FP64 grid-inverse kernel (svla_tok_decode) followed by the q01/q99 un-normalisation (:239-253)."""
        n_tok = 3
        B = generation_outputs.shape[0]
        need = n_tok * self.action_chunk_size
        if generation_outputs.shape[1] < need:
            raise ValueError(f"decode_actions_device needs {need} generated ids per row, got {generation_outputs.shape[1]}")
        ids = generation_outputs[:, :need].reshape(-1, n_tok).contiguous()
        normalized = self.action_tokenizer.decode_ids(ids)                       # (B*chunk, 7) float64 on the device
        if unnorm_key is None:
            logger.warning("unnorm_key None is not in statistics, use next one")
            unnorm_key = next(iter(self.statistics.keys()))
        st = self.statistics[unnorm_key]["action"]
        dev = normalized.device
        hi = torch.tensor(st["q99"], dtype=torch.float64, device=dev)
        lo = torch.tensor(st["q01"], dtype=torch.float64, device=dev)
        mask = torch.tensor(np.array(st.get("mask", np.ones(len(st["q01"]))), dtype=bool), device=dev)
        acts = torch.where(mask, 0.5 * (normalized + 1) * (hi - lo) + lo, normalized)
        return {"actions": acts.view(B, -1, 7), "action_ids": ids.view(B, -1, n_tok)}

    def decode_actions_batch(self, generation_outputs: torch.Tensor, unnorm_key: Optional[str] = None):
        """All rows: (B, >= 3*chunk) ids -> {'actions': (B, chunk, 7), 'action_ids': (B, chunk, 3)}"""
        n_tok = 3
        B = generation_outputs.shape[0]
        ids = generation_outputs[:, : n_tok * self.action_chunk_size].detach().cpu().long().numpy().reshape(-1, n_tok)
        normalized = self.action_tokenizer.decode_token_ids_to_actions(ids)
        acts = self._unnormalize(normalized, unnorm_key)
        return {"actions": acts.reshape(B, -1, 7), "action_ids": ids.reshape(B, -1, n_tok)}
