"""CPU: the LoRA parameter arena of the fine-tune step (config #5) -- target enumeration against the reference's PEFT target list,
flat layout, merge, the adapter chain rule against autograd through the oracle, and the single gradient all-reduce over gloo."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spatialvla_b200.configs import get_config_dict
from spatialvla_b200.lora import LoRAArena, lora_numel, lora_target_keys
from spatialvla_b200.weights import state_dict_spec, synth_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_target_enumeration_matches_reference_recipe():
    """r = 32 on q,k,v,o,gate,up,down + SigLIP q,k,v,out_proj,fc1,fc2 + projector + Ego3D linears (train/spatialvla_finetune.py:262-270,
    scripts/spatialvla_4b_finetune/finetune_lora.sh) = the ~59.1 M trainable parameters of SURVEY.md §8d."""
    spec = state_dict_spec(get_config_dict("4b-224"))
    keys = lora_target_keys(spec, "linear")
    names = {k for k, _, _ in keys}
    assert "language_model.model.layers.25.mlp.down_proj.weight" in names
    assert "vision_tower.vision_model.encoder.layers.26.self_attn.out_proj.weight" in names
    assert "multi_modal_projector.linear.weight" in names
    assert "position_embedding_3d.position_embedding_head.0.weight" in names and "position_embedding_3d.position_embedding_head.3.weight" in names
    assert not any("layernorm" in k or "embed_tokens" in k or "lm_head" in k for k in names)
    n = lora_numel(spec, 32, "linear")
    assert n == 59_184_512 and 59.0e6 < n < 59.3e6
    assert lora_numel(spec, 32, "linear+emb") == n + 32 * (8194 + 2304)
    assert lora_numel(spec, 32, "linear+emb+h") == n + 32 * (8194 + 2304) + 32 * (265347 + 2304)
    with pytest.raises(ValueError):
        lora_target_keys(spec, "everything")


def test_arena_layout_merge_and_chain_rule():
    from oracle import model_ref as R
    cfg = get_config_dict("tiny")
    spec = state_dict_spec(cfg)
    arena = LoRAArena(spec, r=4, alpha=8.0, seed=1)
    assert arena.param.is_contiguous() and arena.grad.is_contiguous() and arena.numel() == lora_numel(spec, 4)
    k = "language_model.model.layers.2.self_attn.q_proj.weight"
    assert arena.A[k].shape == (4, spec[k][1]) and arena.B[k].shape == (spec[k][0], 4)
    assert float(arena.B[k].abs().max()) == 0.0 and 0.15 < float(arena.A[k].std()) < 0.35          # B = 0, A ~ N(0, 1/r)
    arena.A[k].mul_(2.0)                                  # views alias the flat buffer
    assert arena.A[k].data_ptr() >= arena.param.data_ptr() and float(arena.param.abs().sum()) > 0
    sd = synth_state_dict(cfg, seed=0)
    merged = arena.merged_state_dict(sd)
    assert all(torch.equal(merged[kk], sd[kk]) for kk, _, _ in arena.keys)                         # B = 0: merge is the identity
    g = torch.Generator().manual_seed(2)
    arena.B[k].copy_(torch.randn(arena.B[k].shape, generator=g) * 0.05)
    merged = arena.merged_state_dict(sd)
    assert torch.allclose(merged[k], sd[k] + 2.0 * arena.B[k] @ arena.A[k], atol=1e-6) and not torch.equal(merged[k], sd[k])
    # chain rule: gradients of the adapter from the full-weight gradient of the MERGED model == autograd on explicit A, B
    import numpy as np
    gd = np.load(os.path.join(ROOT, "tests", "golden", "tiny_model_train.npz"))
    ids, tt, labels = (torch.from_numpy(gd[n]) for n in ("input_ids", "token_type_ids", "labels"))
    px, K = torch.from_numpy(gd["pixel_u8"]).float() / 255.0, torch.from_numpy(gd["intrinsic"])
    ones = torch.ones_like(ids)
    _, gw = R.loss_and_grads_ref(merged, cfg, ids, px, K, labels, (k,), token_type_ids=tt, attention_mask=ones)
    arena.zero_grad()
    arena.accumulate_from_weight_grad(k, gw[k])
    A = arena.A[k].clone().requires_grad_(True)
    B = arena.B[k].clone().requires_grad_(True)
    sd2 = dict(sd)
    sd2[k] = sd[k] + arena.scale * (B @ A)
    with torch.enable_grad():
        loss, _, _, _ = R._forward_loss(sd2, cfg, ids, px, K, labels, tt, ones, None, -100, 0, None)
        loss.backward()
    assert (arena.gA[k] - A.grad).abs().max() < 1e-5 * max(1.0, float(A.grad.abs().max()))
    assert (arena.gB[k] - B.grad).abs().max() < 1e-5 * max(1.0, float(B.grad.abs().max()))
    assert float(arena.grad.abs().sum()) > 0 and float(arena.gA["multi_modal_projector.linear.weight"].abs().sum()) == 0.0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spatialvla_b200.parallel import allreduce_gradients
    arena = LoRAArena(state_dict_spec(get_config_dict("tiny")), r=4, alpha=8.0, seed=0)       # same seed: identical replicas
    arena.grad.copy_(torch.arange(arena.numel(), dtype=torch.float32) * (rank + 1))
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda t, *a, **k: (calls.append(t.numel()), real(t, *a, **k))[1]
    n = allreduce_gradients(arena)
    dist.all_reduce = real
    if rank == 0:
        q.put((n, calls, arena.grad[:5].tolist(), float(arena.grad[-1]), float(arena.param.abs().sum())))
    dist.destroy_process_group()


def test_gradient_allreduce_is_one_collective_over_the_flat_arena():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n, calls, head, last, psum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    arena = LoRAArena(state_dict_spec(get_config_dict("tiny")), r=4, alpha=8.0, seed=0)
    assert n == arena.numel() and calls == [arena.numel()]                  # exactly one all-reduce, over the whole buffer
    assert head == [0.0, 1.5, 3.0, 4.5, 6.0] and last == (arena.numel() - 1) * 1.5      # mean of ranks: (1 + 2) / 2
    assert abs(psum - float(arena.param.abs().sum())) < 1e-3                # parameters untouched


def test_merged_adapters_serve_through_the_unchanged_engine():
    """Fine-tuned model at inference: adapters folded into the base weights (PEFT merge_and_unload) run through the normal
    predict_action path; the fp32 oracle on the same merged weights is the reference."""
    from oracle import model_ref as R
    from oracle.gen_golden import tiny_inputs
    from oracle.ops_ref import RefOps
    from spatialvla_b200.engine import SpatialVLAEngine
    cfg, px_u8, ids, K = tiny_inputs()
    px = px_u8.float() / 255.0
    sd = synth_state_dict(cfg, seed=0)
    arena = LoRAArena(state_dict_spec(cfg), r=4, alpha=8.0, seed=3)
    g = torch.Generator().manual_seed(4)
    for k, _, _ in arena.keys:
        arena.B[k].copy_(torch.randn(arena.B[k].shape, generator=g) * 0.02)
    merged = arena.merged_state_dict(sd)
    eng = SpatialVLAEngine(cfg, merged, RefOps())
    with torch.no_grad():
        toks, logits = eng.generate_actions(ids, px, K, 4, return_logits=True)
    ref_toks, ref_logits = R.predict_action_ref(merged, cfg, ids, px, K, 4, force_head=eng.last_router_head)
    base_toks, base_logits = R.predict_action_ref(sd, cfg, ids, px, K, 4, force_head=eng.last_router_head)
    assert (logits - ref_logits).abs().max() < 6e-2
    assert (ref_logits - base_logits).abs().max() > 0.2          # the adapters changed the model well beyond the comparison tolerance


def test_adamw_restatement_matches_torch_optimizer():
    """oracle/ops_ref.RefOps.adamw_step (the checker of the svla_adamw_step kernel) against torch.optim.AdamW -- the optimizer the
    reference's HF Trainer runs on the adapters (lr 5e-4, scripts/spatialvla_4b_finetune/finetune_lora.sh:26-29) -- over several steps."""
    from oracle.ops_ref import RefOps
    ops = RefOps()
    for wd in (0.0, 0.01):
        g = torch.Generator().manual_seed(7)
        p_ref = torch.nn.Parameter(torch.randn(1001, generator=g))
        opt = torch.optim.AdamW([p_ref], lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
        p, m, v = p_ref.detach().clone(), torch.zeros(1001), torch.zeros(1001)
        for step in range(1, 5):
            gr = torch.randn(1001, generator=g) * 0.1 * step
            p_ref.grad = gr.clone()
            opt.step()
            ops.adamw_step(p, gr, m, v, lr=5e-4, weight_decay=wd, step=step)
            assert (p - p_ref.detach()).abs().max() < 1e-6, (wd, step)
        st = opt.state[p_ref]
        assert (m - st["exp_avg"]).abs().max() < 1e-7 and (v - st["exp_avg_sq"]).abs().max() < 1e-8
    # grad_scale = the 1 / world-size (or clipping) factor applied on the fly
    p1, m1, v1 = torch.ones(8), torch.zeros(8), torch.zeros(8)
    p2, m2, v2 = torch.ones(8), torch.zeros(8), torch.zeros(8)
    gr = torch.arange(8.0)
    ops.adamw_step(p1, gr, m1, v1, lr=1e-2, step=1, grad_scale=0.25)
    ops.adamw_step(p2, gr * 0.25, m2, v2, lr=1e-2, step=1)
    assert torch.equal(p1, p2) and torch.equal(m1, m2)
