"""Stand-ins for the HF collaborators of the processor (no hub access): tokenizer + image processor."""
import numpy as np
import torch


class FakeTokenizer:
    """Whitespace tokenizer with the HF surface the processor / action tokenizer touch."""
    bos_token, eos_token = "<bos>", "<eos>"
    model_input_names = ["input_ids", "attention_mask"]

    def __init__(self, base=257152):
        self.vocab = {"<pad>": 0, "<eos>": 1, "<bos>": 2, "\n": 108}
        self.base = base
        self.added = {}
        self.init_kwargs = {}

    @property
    def vocab_size(self):
        return self.base

    def __len__(self):
        return self.base + len(self.added)

    def add_special_tokens(self, d):
        return self.add_tokens(d.get("additional_special_tokens", []))

    def add_tokens(self, toks, special_tokens=False):
        n = 0
        for t in toks:
            t = str(t)
            if t not in self.added:
                self.added[t] = self.base + len(self.added)
                n += 1
        return n

    def convert_tokens_to_ids(self, tok):
        if isinstance(tok, (list, tuple, np.ndarray)):
            return [self.convert_tokens_to_ids(t) for t in tok]
        tok = str(tok)
        if tok in self.added:
            return self.added[tok]
        return self.vocab.get(tok, 3 + (hash(tok) % 1000))

    def _encode(self, s):
        import re
        parts = re.findall(r"<[^<>\s]+>|\n|[^\s<]+", s)
        return [self.convert_tokens_to_ids(p) for p in parts]

    def __call__(self, strings, text_pair=None, return_token_type_ids=False, return_tensors="pt", **kw):
        rows, tts = [], []
        for i, s in enumerate(strings):
            a = self._encode(s)
            b = self._encode(text_pair[i]) if text_pair is not None else []
            rows.append(a + b)
            tts.append([0] * len(a) + [1] * len(b))
        L = max(len(r) for r in rows)
        ids = torch.tensor([r + [0] * (L - len(r)) for r in rows])
        out = {"input_ids": ids, "attention_mask": torch.tensor([[1] * len(r) + [0] * (L - len(r)) for r in rows])}
        if return_token_type_ids:
            out["token_type_ids"] = torch.tensor([t + [0] * (L - len(t)) for t in tts])
        return out


class FakeImageProcessor:
    image_seq_length = 256
    size = {"height": 224, "width": 224}
    model_input_names = ["pixel_values"]

    def __call__(self, images, return_tensors="pt", **kw):
        out = []
        for im in images:
            a = np.asarray(im, dtype=np.float32)
            if a.ndim == 3 and a.shape[-1] == 3:
                a = a.transpose(2, 0, 1)
            out.append(torch.from_numpy(a / 255.0 if a.max() > 1.5 else a))
        return {"pixel_values": torch.stack(out)}
