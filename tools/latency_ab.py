"""p50 latency of predict_action at batch 1 (second half of the BASELINE metric), PDL decode chain vs the persistent single-launch
decode kernel (svla_decode_mega_step), alternated A/B/A/B inside ONE process on the same box: same weights, same inputs, CUDA-graph
replay, wall clock around generate_actions with a device synchronize on both sides (what bench.py's latency leg measures).
Usage: python tools/latency_ab.py [--iters 40]"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from bench import N_NEW, synth_inputs
from spatialvla_b200 import get_config_dict
from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
from spatialvla_b200.weights import synth_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--batch", type=int, default=1)
args = ap.parse_args()
dev = "cuda:0"
cfg = get_config_dict("4b-224")
sd = synth_state_dict(cfg, seed=0, device=dev, on_device_rng=True, dtype=torch.bfloat16)
model = SpatialVLAForConditionalGeneration(cfg, sd, device=dev)
del sd
torch.cuda.empty_cache()
eng = model.engine
px, ids, K = synth_inputs(cfg, args.batch, device=dev)


def measure(small):
    eng.mega_decode = small
    eng._graphs = {}                      # graphs bake the decode path: re-capture for this arm
    ts = []
    toks = None
    for i in range(args.iters + 5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        toks = eng.generate_actions(ids, px, K, N_NEW)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = ts[5:]
    return {"mega_decode": small, "p50_ms": round(statistics.median(ts), 3), "p10_ms": round(sorted(ts)[len(ts) // 10], 3),
            "p90_ms": round(sorted(ts)[len(ts) * 9 // 10], 3)}, toks


res = []
ref = None
for small in (False, True, False, True):
    r, toks = measure(small)
    ref = toks if ref is None else ref
    r["tokens_equal_first_arm"] = bool(torch.equal(toks, ref))
    res.append(r)
    print(json.dumps(r), flush=True)
