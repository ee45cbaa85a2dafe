"""`SpatialActionTokenizer` with the reference's interface (model/action_tokenizer.py:249-430); the grid lookup and
its inverse run in the FP64 CUDA kernels behind `svla_tok_encode_host` / `svla_tok_decode_host` (HOST buffers in,
HOST buffers out -- the call a dataset / processor makes) or, for device-resident batches, `encode_ids` /
`decode_ids` on torch CUDA tensors.  Token strings `<ACTION%05d>` only appear at the API edge.

Bin edges are configuration (6 small float64 arrays), computed on the host exactly like the reference's
`get_bin_policy` (scipy.stats.norm cdf/ppf, model/action_tokenizer.py:343-370)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib as L

ACTION_TOKEN = "<ACTION{:05d}>"
_TRANS_KEYS = ("theta_bins", "phi_bins", "r_bins")
_ROT_KEYS = ("roll_bins", "pitch_bins", "yaw_bins")


class _SubTokenizer:
    """Attribute surface of the reference's Translation/Rotation/Gripper tokenizers that other code reads
    (train/monkey_patch.py:273-297): token_start_idx, token_end_idx, vocab_size, token_array."""

    def __init__(self, tokenizer, n, array_begin_idx):
        self.tokenizer = tokenizer
        self.array_begin_idx = array_begin_idx
        self._vocab_size = n
        self.token_array = np.array([ACTION_TOKEN.format(i + array_begin_idx) for i in range(n)])
        self.tokenizer.add_tokens(list(self.token_array), special_tokens=True)
        self.token_start_idx = self.tokenizer.convert_tokens_to_ids(self.token_array[0])
        self.token_end_idx = self.tokenizer.convert_tokens_to_ids(self.token_array[-1])

    @property
    def vocab_size(self):
        return self._vocab_size


def edge_trig_table(theta_edges, phi_edges):
    """Table behind the exact angular binning of svla_tok_encode (csrc/tokenizer.cu): for every interior theta / phi edge e the
    reference's test `fl(atan2(a, b)) >= e` (np.digitize, model/action_tokenizer.py:115-118) is, for a correctly rounded atan2,
    `atan2(a, b) >= m` with m = (pred(e) + e) / 2, decided on the device as the sign of a cos m - b sin m.  cos m / sin m are
    evaluated here with 200-bit arithmetic and stored as double-double pairs.
    Returns (float64 [n, 4] rows (cos_hi, cos_lo, sin_hi, sin_lo), #phi edges <= 0, #phi edges < 0)."""
    import mpmath as mp
    rows = []
    with mp.workprec(200):
        for e in list(theta_edges) + list(phi_edges):
            e = float(e)
            m = (mp.mpf(float(np.nextafter(e, -np.inf))) + mp.mpf(e)) / 2
            c, s_ = mp.cos(m), mp.sin(m)
            ch, sh = float(c), float(s_)
            rows.append((ch, float(c - mp.mpf(ch)), sh, float(s_ - mp.mpf(sh))))
    phi = np.asarray(phi_edges, dtype=np.float64)
    tab = np.ascontiguousarray(np.asarray(rows, dtype=np.float64).reshape(-1, 4))
    return tab, int((phi <= 0).sum()), int((phi < 0).sum())


class SpatialActionTokenizer:
    range_bins = {
        "translation": {"theta_bins": (0.0, np.pi), "phi_bins": (-np.pi, np.pi), "r_bins": (0.0, np.sqrt(3))},
        "rotation": {"roll_bins": (-1.0, 1.0), "pitch_bins": (-1.0, 1.0), "yaw_bins": (-1.0, 1.0)},
    }

    def __init__(self, tokenizer, num_bins: Dict, gs_params: Dict = None, bin_policy: Dict = None,
                 use_spherical: bool = True, min_sigma: float = 0.0, min_action: float = -1.0, max_action: float = 1.0):
        self.tokenizer = tokenizer
        self.min_action, self.max_action = min_action, max_action
        self.num_bins = num_bins
        self.min_sigma = min_sigma
        self.use_spherical = use_spherical
        self.bin_policy = bin_policy if bin_policy else self.get_bin_policy(gs_params, self.min_sigma)
        nt, nr = num_bins["translation"], num_bins["rotation"]
        n_trans = nt["theta_bins"] * nt["phi_bins"] * nt["r_bins"]
        n_rot = nr["roll_bins"] * nr["pitch_bins"] * nr["yaw_bins"]
        self.translation_tokenizer = _SubTokenizer(tokenizer, n_trans, 0)
        self.rotation_tokenizer = _SubTokenizer(tokenizer, n_rot, n_trans)
        self.gripper_tokenizer = _SubTokenizer(tokenizer, num_bins["gripper"], n_trans + n_rot)
        self._vocab_size = n_trans + n_rot + num_bins["gripper"]
        self.token_array = np.concatenate([self.translation_tokenizer.token_array, self.rotation_tokenizer.token_array,
                                           self.gripper_tokenizer.token_array])
        self._refresh_edges()

    # ---- configuration
    def _refresh_edges(self):
        pol = self.bin_policy
        arrs = [np.asarray(pol["translation"][k], dtype=np.float64) for k in _TRANS_KEYS] + \
               [np.asarray(pol["rotation"][k], dtype=np.float64) for k in _ROT_KEYS]
        nb = [self.num_bins["translation"][k] for k in _TRANS_KEYS] + [self.num_bins["rotation"][k] for k in _ROT_KEYS]
        for a, n in zip(arrs, nb):
            if a.shape[0] != n + 1:
                raise ValueError(f"bin policy has {a.shape[0]} edges for {n} bins")
        self._edges = np.ascontiguousarray(np.concatenate(arrs))
        self._nbins = (C.c_int32 * 7)(*nb, int(self.num_bins["gripper"]))
        self._nbins_list = nb + [int(self.num_bins["gripper"])]
        self._dev_edges = {}
        self._trig, self._phi_nonpos, self._phi_neg = edge_trig_table(arrs[0][1:-1], arrs[1][1:-1])
        # (sin, cos) of the theta / phi bin centres: the only angles decode evaluates (model/action_tokenizer.py:99-103,129-135)
        cen = [0.5 * (a[:-1] + a[1:]) for a in arrs[:2]]
        self._ctrig = np.ascontiguousarray(np.concatenate([np.stack([np.sin(c), np.cos(c)], 1) for c in cen]).astype(np.float64))

    @property
    def vocab_size(self) -> int:
        return self._vocab_size

    @property
    def action_token_begin_idx(self) -> int:
        return self.translation_tokenizer.token_start_idx

    def get_bin_policy(self, gs_params=None, min_sigma=0.0):
        """model/action_tokenizer.py:343-370: Gaussian-quantile edges clipped to the axis range, or uniform."""
        from scipy.stats import norm
        pol = {"translation": {}, "rotation": {}}
        for bt, axes in self.range_bins.items():
            for bk, (lo, hi) in axes.items():
                n = self.num_bins[bt][bk]
                if gs_params is None:
                    pol[bt][bk] = np.linspace(lo, hi, n + 1)
                else:
                    g = gs_params[bk.split("_")[0].lower()]
                    mu, sigma = g["mu"], max(g["sigma"], min_sigma)
                    prob = np.linspace(norm.cdf(lo, loc=mu, scale=sigma), norm.cdf(hi, loc=mu, scale=sigma), n + 1)
                    pol[bt][bk] = np.clip(norm.ppf(prob, loc=mu, scale=sigma), lo, hi).tolist()
        return pol

    # ---- host-buffer API (reference signatures)
    def encode_local_ids(self, action: np.ndarray) -> np.ndarray:
        """(n,7)|(7,) float -> (n,3) int32 local ids in [0, vocab). Runs svla_tok_encode_host."""
        action = np.asarray(action)
        if action.ndim == 1:
            assert action.shape[0] == 7, f"action dim mismatch, got action shape: {action.shape}"
            action = action.reshape(1, 7)
        assert action.shape[1] == 7, f"action dim mismatch, got action shape: {action.shape}"
        a = np.ascontiguousarray(action, dtype=np.float64)
        ids = np.empty((a.shape[0], 3), dtype=np.int32)
        lib = L.load_library()
        L.check(lib.svla_tok_encode_host(a.ctypes.data, self._edges.ctypes.data, C.cast(self._nbins, C.c_void_p),
                                         ids.ctypes.data, a.shape[0], float(self.min_action), float(self.max_action),
                                         int(self.use_spherical), self._trig.ctypes.data, self._phi_nonpos, self._phi_neg),
                "svla_tok_encode_host")
        return ids

    def __call__(self, action: np.ndarray) -> np.ndarray:
        """Discretize continuous actions (n,7) to token strings (n,3) (model/action_tokenizer.py:305-319)."""
        return self.token_array[self.encode_local_ids(action)]

    def decode_token_ids_to_actions(self, action_token_ids: np.ndarray) -> np.ndarray:
        """(n,3)|(3,) global token ids -> (n,7) float64 actions (model/action_tokenizer.py:321-333)."""
        ids = np.asarray(action_token_ids)
        if ids.ndim == 1:
            assert ids.shape[0] == 3, f"action token id numbers mismatich, need 3 got {ids.shape[0]}"
            ids = ids.reshape(1, 3)
        assert ids.shape[1] == 3, f"token id numbers mismatich, need 3 got {ids.shape[1]}"
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        out = np.empty((ids.shape[0], 7), dtype=np.float64)
        lib = L.load_library()
        L.check(lib.svla_tok_decode_host(ids.ctypes.data, self._edges.ctypes.data, C.cast(self._nbins, C.c_void_p),
                                         int(self.action_token_begin_idx), out.ctypes.data, ids.shape[0],
                                         int(self.use_spherical), self._ctrig.ctypes.data), "svla_tok_decode_host")
        return out

    # ---- device-resident API (batched serving / training data path)
    def _edges_on(self, device):
        import torch
        key = str(device)
        if key not in self._dev_edges:
            self._dev_edges[key] = (torch.from_numpy(self._edges).to(device), torch.from_numpy(self._trig).to(device),
                                    torch.from_numpy(self._ctrig).to(device))
        return self._dev_edges[key]

    def encode_ids(self, actions):
        """torch float64 CUDA tensor (n,7) -> int64 (n,3) GLOBAL token ids, asynchronous on the current stream."""
        import torch
        from .ops import CudaOps
        ops = CudaOps(actions.device)
        a = actions.to(torch.float64).contiguous()
        ids = torch.empty((a.shape[0], 3), dtype=torch.int32, device=a.device)
        edges, trig, _ = self._edges_on(a.device)
        ops.tok_encode(a, edges, self._nbins_list, ids, min_action=self.min_action, max_action=self.max_action,
                       use_spherical=self.use_spherical, trig=trig, phi_nonpos=self._phi_nonpos, phi_neg=self._phi_neg)
        return ids.to(torch.int64) + self.action_token_begin_idx

    def decode_ids(self, ids):
        """torch int64 CUDA tensor (n,3) of global ids -> float64 (n,7)"""
        import torch
        from .ops import CudaOps
        ops = CudaOps(ids.device)
        i = ids.to(torch.int64).contiguous()
        out = torch.empty((i.shape[0], 7), dtype=torch.float64, device=i.device)
        edges, _, ctrig = self._edges_on(i.device)
        ops.tok_decode(i, edges, self._nbins_list, self.action_token_begin_idx, out, use_spherical=self.use_spherical,
                       center_trig=ctrig)
        return out

    # ---- fine-tune-time re-gridding (host-side one-off, mirrors :372-430)
    def get_norm_meshgrid(self, bin_policy):
        grids = []
        policy = {k1: {k2: np.array(v2) for k2, v2 in v1.items()} for k1, v1 in bin_policy.items()}
        for bin_type in self.range_bins.keys():
            bounds = []
            for bin_key in self.range_bins[bin_type].keys():
                minb, maxb = self.range_bins[bin_type][bin_key]
                edges = policy[bin_type][bin_key]
                centre = np.concatenate([np.array([minb]), (edges[:-1] + edges[1:]) / 2, np.array([maxb])])
                bounds.append((centre - minb) / (maxb - minb))
            gx, gy, gz = np.meshgrid(*bounds)
            grids += [np.stack([gx, gy, gz], -1).reshape(-1, 3)]
        return grids[0], grids[1]

    @staticmethod
    def adaption_plan(grid0, grid1, dims):
        """Geometry of the re-gridding, independent of the embedding width: `griddata(grid0, values, grid1, 'linear')` is Delaunay
        interpolation -- NOT trilinear: Qhull splits the (degenerate) cells of the padded grid into tetrahedra -- so the plan takes
        the triangulation from the same Qhull call (scipy.spatial.Delaunay, what griddata builds internally), locates every
        interior target point and returns its 4 source rows and barycentric weights.  Point p of grid0 carries the value of padded
        cell unravel(p, (m+2, n+2, k+2)) exactly as the reference pairs them (its meshgrid is 'xy'-ordered while the values are
        'ij'-ordered, model/action_tokenizer.py:372-388,402-405 -- reproduced, not fixed); replicate padding = clamped indices.
        -> (rows int32 [m*n*k, 4] embedding rows of the block, -1 where the target lies outside the hull (NaN in the reference),
            weights float64 [m*n*k, 4])."""
        from scipy.spatial import Delaunay
        m, n, k = dims
        tri = Delaunay(np.asarray(grid0, dtype=np.float64))
        ii, jj, ll = np.meshgrid(np.arange(1, m + 1), np.arange(1, n + 1), np.arange(1, k + 1), indexing="ij")
        tgt = ((ii * (n + 2) + jj) * (k + 2) + ll).reshape(-1)                 # interior targets, in output-row order
        x = np.asarray(grid1, dtype=np.float64)[tgt]
        simp = tri.find_simplex(x)
        ok = simp >= 0
        T = tri.transform[np.where(ok, simp, 0)]
        b = np.einsum("nij,nj->ni", T[:, :3], x - T[:, 3])
        w = np.concatenate([b, 1.0 - b.sum(1, keepdims=True)], 1)
        verts = tri.simplices[np.where(ok, simp, 0)]                           # [N, 4] point indices of grid0
        pi, pj, pl = np.unravel_index(verts, (m + 2, n + 2, k + 2))
        rows = (np.clip(pi - 1, 0, m - 1) * n + np.clip(pj - 1, 0, n - 1)) * k + np.clip(pl - 1, 0, k - 1)
        rows = np.where(ok[:, None], rows, -1).astype(np.int32)
        return np.ascontiguousarray(rows), np.ascontiguousarray(w)

    def spatial_embedding_adaption(self, gs_params, embeddings, min_sigma=0.0, adpt_feature=False, ops=None):
        """Re-grid the tokenizer to new Gaussians and (optionally) re-sample the spatial embeddings by scattered linear
        interpolation, as model/action_tokenizer.py:390-430 does.  With `ops` (CudaOps; implied for CUDA embeddings) the 8 192 x E
        re-sampling runs on the GPU: the triangulation / point location (E-independent, `adaption_plan`) stays on the host, the
        4-row barycentric gather over the embedding table is the `svla_barycentric_gather` kernel.  Without it: scipy griddata on
        the host, call for call like the reference."""
        import torch
        new_policy = self.get_bin_policy(gs_params, min_sigma=min_sigma)
        g0t, g0r = self.get_norm_meshgrid(self.bin_policy)
        g1t, g1r = self.get_norm_meshgrid(new_policy)
        self.bin_policy, self.min_sigma = new_policy, min_sigma
        self._refresh_edges()
        if not adpt_feature:
            return
        emb = embeddings.weight.data
        E = emb.shape[1]
        off = 0
        if ops is None and emb.is_cuda:
            from .ops import CudaOps
            ops = CudaOps(emb.device)
        if ops is not None:
            for (grid0, grid1, keys, bt) in ((g0t, g1t, _TRANS_KEYS, "translation"), (g0r, g1r, _ROT_KEYS, "rotation")):
                dims = tuple(self.num_bins[bt][kk] for kk in keys)
                N = dims[0] * dims[1] * dims[2]
                rows, w = self.adaption_plan(grid0, grid1, dims)
                src = emb[off:off + N].to(device=ops.device, dtype=torch.float32).contiguous()
                out = ops.empty((N, E), torch.float32)
                ops.barycentric_gather(src, torch.from_numpy(rows).to(ops.device), torch.from_numpy(w).to(ops.device), out)
                emb[off:off + N] = out.to(device=emb.device, dtype=emb.dtype)
                off += N
            return
        from scipy.interpolate import griddata
        for (grid0, grid1, keys, bt) in ((g0t, g1t, _TRANS_KEYS, "translation"), (g0r, g1r, _ROT_KEYS, "rotation")):
            m, n, k = (self.num_bins[bt][kk] for kk in keys)
            N = m * n * k
            blk = emb[off:off + N].reshape(m, n, k, -1).permute(3, 0, 1, 2)
            pad = torch.nn.functional.pad(blk, (1, 1, 1, 1, 1, 1), "replicate").permute(1, 2, 3, 0).reshape(-1, E)
            ad = griddata(grid0, pad.float().cpu().numpy(), grid1, method="linear")
            ad = ad.reshape(m + 2, n + 2, k + 2, E)[1:-1, 1:-1, 1:-1]
            emb[off:off + N] = torch.from_numpy(ad.reshape(-1, E)).to(emb.dtype).to(emb.device)
            off += N
