"""The LoRA fine-tune step (BASELINE.json config #5): forward with un-merged adapters + full backward +
clip + AdamW, host orchestration (spatialvla_b200/training.py) run over the torch op re-statements on the CPU and -- `-m gpu` --
over the CUDA kernels, against the autograd oracle (oracle/model_ref.loss_and_grads_ref, pinned on loss.backward() of the LIVE
reference by tests/golden/tiny_model_train_grads.npz).  The oracle differentiates the MERGED weights W + (alpha/r) B A; the
adapter gradients follow by the chain rule gA = s B^T dW, gB = s dW A^T (checked against autograd in test_oracle_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref as R
from oracle.gen_golden import train_inputs
from oracle.ops_ref import RefOps
from spatialvla_b200.engine import SpatialVLAEngine
from spatialvla_b200.training import LoRATrainer
from spatialvla_b200.weights import synth_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _setup(ops, seed=0, window=None):
    cfg, px_u8, ids, tt, labels, K = train_inputs()
    if window:          # sliding-window layers active: window << the 270-token training sequence
        cfg["text_config"]["sliding_window"] = window
    sd = synth_state_dict(cfg, seed=0)
    eng = SpatialVLAEngine(cfg, sd, ops)
    tr = LoRATrainer(eng, r=32, alpha=32.0, lr=1e-3, seed=seed)
    tr.lay.randomize_B(std=0.02, seed=1)
    return cfg, px_u8.float() / 255.0, ids, tt, labels, K, sd, eng, tr


def _oracle_adapter_grads(tr, sd, cfg, ids, px, K, labels, tt, am, head):
    lay = tr.lay
    # the step's operands are the bf16 roundings of the fp32 master adapters: the oracle merges exactly those
    merged = dict(sd)
    Ab = {k: lay.A(k).to(torch.bfloat16).float().cpu() for k, _, _ in lay.keys}
    Bb = {k: lay.Bt(k).to(torch.bfloat16).float().t().cpu() for k, _, _ in lay.keys}
    active = [a[0] for fl in lay.fused.values() for a in fl.adapters]
    for k in active:
        merged[k] = sd[k].float() + lay.scale * (Bb[k] @ Ab[k])
    loss, dW = R.loss_and_grads_ref(merged, cfg, ids, px, K, labels, tuple(active), token_type_ids=tt, attention_mask=am, force_head=head)
    out = {}
    for k in active:
        out[k] = (lay.scale * (Bb[k].t() @ dW[k]), lay.scale * (dW[k] @ Ab[k].t()).t())       # gA [r, in], gB^T [r, out]
    return float(loss), out


def _check_grads(tr, ref, tol):
    worst = {}
    for fl in tr.lay.fused.values():
        for (k, fin, fout, _, _) in fl.adapters:
            ga, gb = tr.lay.gA(k).float().cpu(), tr.lay.gBt(k).float().cpu()
            ra, rb = ref[k]
            ea = float((ga - ra).abs().max()) / max(float(ra.abs().max()), 1e-12)
            eb = float((gb - rb).abs().max()) / max(float(rb.abs().max()), 1e-12)
            worst[k] = (ea, eb)
    bad = {k: v for k, v in worst.items() if max(v) > tol}
    assert not bad, f"{len(bad)} of {len(worst)} adapters off: " + str(sorted(bad.items(), key=lambda kv: -max(kv[1]))[:6])
    return worst


@pytest.mark.parametrize("mask", ["prefix_lm", "causal", "prefix_lm_window48"])
def test_lora_step_host_logic_gradients_match_autograd_oracle(mask):
    cfg, px, ids, tt, labels, K, sd, eng, tr = _setup(RefOps(), window=48 if mask.endswith("window48") else None)
    B, L = ids.shape
    am = torch.ones(B, L, dtype=torch.int64) if mask.startswith("prefix_lm") else None
    summary = tr.forward_backward(ids, px, K, labels, token_type_ids=tt, attention_mask=am)
    ref_loss, ref = _oracle_adapter_grads(tr, sd, cfg, ids, px, K, labels, tt, am, eng.last_router_head)
    assert abs(float(summary[0]) - ref_loss) < 2e-2, (float(summary[0]), ref_loss)
    worst = _check_grads(tr, ref, 6e-2)
    # the adapters PEFT creates on frozen ZoeDepth modules keep exactly zero gradient
    active = {a[0] for fl in tr.lay.fused.values() for a in fl.adapters}
    passive = [k for k, _, _ in tr.lay.keys if k not in active]
    assert passive and all(float(tr.lay.gA(k).abs().max()) == 0 and float(tr.lay.gBt(k).abs().max()) == 0 for k in passive)
    assert len(worst) == len(active)


def test_lora_step_with_zero_B_equals_the_base_model_and_updates_only_B():
    """PEFT's init (B = 0): the adapted forward equals the base model's loss (golden of the live reference), gA is exactly zero,
    and one optimizer step moves B only."""
    cfg, px_u8, ids, tt, labels, K = train_inputs()
    sd = synth_state_dict(cfg, seed=0)
    eng = SpatialVLAEngine(cfg, sd, RefOps())
    tr = LoRATrainer(eng, lr=1e-3)
    g = np.load(os.path.join(GOLD, "tiny_model_train.npz"))
    B, L = ids.shape
    batch = {"input_ids": ids, "pixel_values": px_u8.float() / 255.0, "intrinsic": K, "labels": labels, "token_type_ids": tt,
             "attention_mask": torch.ones(B, L, dtype=torch.int64)}
    p0 = tr.lay.param.clone()
    summary = tr.step(batch)
    assert abs(float(summary[0]) - float(g["loss_prefix_lm"])) < 2e-2
    k = "language_model.model.layers.1.self_attn.q_proj.weight"
    assert float(tr.lay.gA(k).abs().max()) == 0 and float(tr.lay.gBt(k).abs().max()) > 0
    a0, b0 = tr.lay.slots[k]
    moved = (tr.lay.param - p0).abs()
    assert float(moved[a0:b0].max()) == 0 and float(moved[b0:b0 + 32 * 1024].max()) > 0
    # clip_grad_norm_(1.0): the applied update equals AdamW on the clipped gradient
    gn = float(tr.lay.grad.norm())
    assert gn > 0 and tr.step_count == 1


@pytest.mark.gpu
@pytest.mark.parametrize("mask", ["prefix_lm", "causal", "causal_window48"])
def test_lora_step_on_gpu_gradients_match_autograd_oracle(mask, cuda_device):
    from spatialvla_b200.ops import CudaOps
    ops = CudaOps(cuda_device)
    cfg, px, ids, tt, labels, K, sd, eng, tr = _setup(ops, window=48 if mask.endswith("window48") else None)
    B, L = ids.shape
    am = torch.ones(B, L, dtype=torch.int64) if mask == "prefix_lm" else None
    n0 = ops.launch_count()
    summary = tr.forward_backward(ids, px, K, labels, token_type_ids=tt, attention_mask=am)
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 > 200              # every FLOP of the step runs in this library's kernels
    ref_loss, ref = _oracle_adapter_grads(tr, sd, cfg, ids, px, K, labels, tt, am, eng.last_router_head)
    assert abs(float(summary[0]) - ref_loss) < 2e-2, (float(summary[0]), ref_loss)
    worst = _check_grads(tr, ref, 6e-2)
    print("lora step gpu: worst adapter errors", sorted(worst.items(), key=lambda kv: -max(kv[1]))[:3])
    # the nine weights whose full gradients the LIVE reference minted (tests/golden/tiny_model_train_grads.npz): with B = 0 the
    # adapter gradient is gB = s dW A^T of the BASE model, so the golden dW pins it directly
    if mask == "prefix_lm":
        from oracle.gen_golden import GRAD_KEYS
        eng2 = SpatialVLAEngine(cfg, sd, ops)
        tr0 = LoRATrainer(eng2)
        tr0.forward_backward(ids, px, K, labels, token_type_ids=tt, attention_mask=am)
        gg = np.load(os.path.join(GOLD, "tiny_model_train_grads.npz"))
        _, dW = R.loss_and_grads_ref(sd, cfg, ids, px, K, labels, GRAD_KEYS, token_type_ids=tt, attention_mask=am,
                                     force_head=eng2.last_router_head)
        for k in GRAD_KEYS:
            if dW[k].dim() != 2:
                continue                         # a norm weight: frozen under LoRA, no adapter
            assert np.abs(dW[k][::3, ::3].numpy() - gg["grad:" + k]).max() < 2e-4 * np.abs(gg["grad:" + k]).max() + 1e-7      # oracle == live reference
            Ab = tr0.lay.A(k).to(torch.bfloat16).float().cpu()
            want = tr0.lay.scale * (dW[k] @ Ab.t()).t()
            got = tr0.lay.gBt(k).float().cpu()
            assert float((got - want).abs().max()) <= 6e-2 * float(want.abs().max()), k


@pytest.mark.gpu
def test_lora_optimizer_step_on_gpu_matches_restatement(cuda_device):
    """sumsq + clip + AdamW kernels on the arena vs the torch re-statement (itself equal to torch.optim.AdamW + clip_grad_norm_)."""
    from spatialvla_b200.ops import CudaOps
    ops, ref = CudaOps(cuda_device), RefOps()
    g = torch.Generator().manual_seed(0)
    n = 1_000_003
    p0, gr = torch.randn(n, generator=g), torch.randn(n, generator=g) * 0.01
    outs = []
    for o, dev in ((ops, cuda_device), (ref, "cpu")):
        p, m, v, s = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(1, device=dev)
        for step in (1, 2):
            s.zero_()
            o.sumsq(gr.to(dev), s)
            o.adamw_step(p, gr.to(dev), m, v, lr=5e-4, step=step, grad_scale=0.125, sumsq=s, max_grad_norm=1.0)
        outs.append((p.cpu(), float(s)))
    assert abs(outs[0][1] - outs[1][1]) < 1e-4 * outs[1][1]
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 2e-6
    assert float((outs[1][0] - p0).abs().max()) > 1e-4


# ------------------------------------------------------------------------------------------------ data-parallel step (gloo, CPU)
def _dp_worker(rank, world, port, q):
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spatialvla_b200.parallel import shard_batch
    torch.set_num_threads(2)
    cfg, px, ids, tt, labels, K, sd, eng, tr = _setup(RefOps())
    eng.force_head = 0                                   # ZoeDepth's router votes over the local batch: pin it for the comparison
    B, L = ids.shape
    batch = {"input_ids": ids, "pixel_values": px, "intrinsic": K, "labels": labels, "token_type_ids": tt,
             "attention_mask": torch.ones(B, L, dtype=torch.int64)}
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda t, *a, **k: (calls.append(t.numel()), real(t, *a, **k))[1]
    tr.step(shard_batch(batch, rank, world))
    dist.all_reduce = real
    q.put((rank, calls, tr.lay.grad.clone(), tr.lay.param.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_step_equals_single_process_step_on_the_whole_batch():
    """World size 2 (gloo): every rank runs forward/backward on its shard, the gradient arena is exchanged in two overlapped
    all-reduces over disjoint slices (Gemma2 segment first), AdamW divides by the world size: summed gradients / 2 equal the
    single-process gradient of the whole batch (each sample labels the same number of tokens), and both replicas end identical."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 34500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, calls0, g0, p0), (_, calls1, g1, p1) = res
    cfg, px, ids, tt, labels, K, sd, eng, tr = _setup(RefOps())
    eng.force_head = 0
    B, L = ids.shape
    tr.forward_backward(ids, px, K, labels, token_type_ids=tt, attention_mask=torch.ones(B, L, dtype=torch.int64))
    n, n1 = tr.lay.numel(), tr.lay.n_language
    assert 0 < n1 < n and calls0 == [n1, n - n1] and calls1 == calls0            # two collectives over disjoint slices of ONE buffer
    assert torch.equal(g0, g1) and torch.equal(p0, p1)                           # replicas stay bit-identical
    ref = tr.lay.grad
    err = float((g0 / 2 - ref).abs().max()) / float(ref.abs().max())
    assert err < 2e-2, err
