"""Config #4 micro-benchmarks (BASELINE.json): SpatialActionTokenizer on 1 M actions and the fused Ego3D encode on
batches of depth maps -- achieved HBM GB/s (algorithmic bytes / CUDA-event time) against MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from spatialvla_b200.ops import CudaOps
from spatialvla_b200.configs import default_intrinsic_224

dev = "cuda:0"
ops = CudaOps(dev)
peak = 6555.8
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 * 2**20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best

g = np.load(os.path.join(ROOT, "tests", "golden", "tokenizer_gauss.npz"))
keys = ("theta_bins", "phi_bins", "r_bins", "roll_bins", "pitch_bins", "yaw_bins")
edges = torch.from_numpy(np.concatenate([g[f"edge_{k}"] for k in keys])).to(dev)
nb = [len(g[f"edge_{k}"]) - 1 for k in keys] + [2]
# the product path (SpatialActionTokenizer.encode_ids / decode_ids): exact angular binning through the edge table, decode through
# the host-tabulated (sin, cos) of the bin centres
from spatialvla_b200.action_tokenizer import edge_trig_table
th_e, ph_e = g["edge_theta_bins"], g["edge_phi_bins"]
trig_np, phi_nonpos, phi_neg = edge_trig_table(th_e[1:-1], ph_e[1:-1])
trig = torch.from_numpy(trig_np).to(dev)
cen = [0.5 * (a[:-1] + a[1:]) for a in (th_e, ph_e)]
ctrig = torch.from_numpy(np.ascontiguousarray(np.concatenate([np.stack([np.sin(c), np.cos(c)], 1) for c in cen]).astype(np.float64))).to(dev)
out = []
for n in (1_000_000, 16_000_000):
    acts = (torch.rand(n, 7, device=dev, dtype=torch.float64) * 2 - 1)
    ids = torch.empty(n, 3, dtype=torch.int32, device=dev)
    ms = timeit(lambda: ops.tok_encode(acts, edges, nb, ids, trig=trig, phi_nonpos=phi_nonpos, phi_neg=phi_neg))
    out.append({"kernel": "svla_tok_encode", "n": n, "ms": round(ms, 4), "GBs": round(n * 68 / ms / 1e6, 1), "frac_of_measured_hbm": round(n * 68 / ms / 1e6 / peak, 3)})
    gid = ids.to(torch.int64) + 257153
    dec = torch.empty(n, 7, dtype=torch.float64, device=dev)
    ms = timeit(lambda: ops.tok_decode(gid, edges, nb, 257153, dec, center_trig=ctrig))
    out.append({"kernel": "svla_tok_decode", "n": n, "ms": round(ms, 4), "GBs": round(n * 80 / ms / 1e6, 1), "frac_of_measured_hbm": round(n * 80 / ms / 1e6 / peak, 3)})
K = torch.tensor(default_intrinsic_224(), device=dev)
for B in (64, 4096):
    depth = torch.rand(B, 384, 384, device=dev) * 4.7 + 0.3
    xyz = torch.empty(B * 256, 12, device=dev); enc = torch.empty(B * 256, 208, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.ego3d_encode(depth, K, xyz, enc, n_freqs=8), iters=5)
    byts = B * (384 * 384 * 4 + 256 * 12 * 4 + 256 * 208 * 2)
    # the 224-crop only depends on source rows / columns 40..343 of the 384-map: bytes the kernel has to fetch at minimum
    need = B * (304 * 304 * 4 + 256 * 12 * 4 + 256 * 208 * 2)
    out.append({"kernel": "svla_ego3d_encode", "maps": B, "ms": round(ms, 4), "GBs": round(byts / ms / 1e6, 1), "frac_of_measured_hbm": round(byts / ms / 1e6 / peak, 3),
                "GBs_crop_window_only": round(need / ms / 1e6, 1), "frac_crop_window_only": round(need / ms / 1e6 / peak, 3)})
for r in out:
    print(json.dumps(r))
