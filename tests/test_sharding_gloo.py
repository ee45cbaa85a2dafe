"""CPU, world_size 2 over gloo: the N>1 inference path shards the observation batch contiguously, runs a full replica
per rank with NO data-path collective, and the gathered tokens equal the single-process result (router head pinned,
since ZoeDepth's vote is over the local batch -- SURVEY.md §8e)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle.gen_golden import tiny_inputs
    from oracle.ops_ref import RefOps
    from spatialvla_b200.engine import SpatialVLAEngine
    from spatialvla_b200.parallel import shard_batch
    from spatialvla_b200.weights import synth_state_dict
    cfg, px_u8, ids, K = tiny_inputs(B=2)
    eng = SpatialVLAEngine(cfg, synth_state_dict(cfg, seed=0), RefOps())
    eng.force_head = 0
    batch = shard_batch({"input_ids": ids, "pixel_values": px_u8.float() / 255.0, "intrinsic": K}, rank, world)
    with torch.no_grad():
        toks = eng.generate_actions(batch["input_ids"], batch["pixel_values"], batch["intrinsic"], 3)
    # timing-style reduction only (max over ranks), exactly what bench.py does: no tensor of the data path is exchanged
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [torch.zeros_like(toks) for _ in range(world)]
    dist.all_gather(gathered, toks)          # test-only gather to compare with the single-process run
    if rank == 0:
        q.put((torch.cat(gathered, 0), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_replicas_match_single_process():
    sys.path.insert(0, ROOT)
    from oracle.gen_golden import tiny_inputs
    from oracle.ops_ref import RefOps
    from spatialvla_b200.engine import SpatialVLAEngine
    from spatialvla_b200.weights import synth_state_dict
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    toks2, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, px_u8, ids, K = tiny_inputs(B=2)
    eng = SpatialVLAEngine(cfg, synth_state_dict(cfg, seed=0), RefOps())
    eng.force_head = 0
    with torch.no_grad():
        toks1 = eng.generate_actions(ids, px_u8.float() / 255.0, K, 3)
    assert torch.equal(toks1, toks2) and tmax == 2.0


def _loss_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import numpy as np
    from oracle.ops_ref import RefOps
    from spatialvla_b200.configs import get_config_dict
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    from spatialvla_b200.parallel import reduce_loss, shard_batch
    from spatialvla_b200.weights import synth_state_dict
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_model_train.npz"))
    cfg = get_config_dict("tiny")
    batch = {"input_ids": torch.from_numpy(g["input_ids"]), "pixel_values": torch.from_numpy(g["pixel_u8"]).float() / 255.0,
             "intrinsic": torch.from_numpy(g["intrinsic"]), "labels": torch.from_numpy(g["labels"]),
             "token_type_ids": torch.from_numpy(g["token_type_ids"])}
    batch["labels"][1, -3:] = -100            # ranks label different numbers of tokens: the reduction must weight by rows
    mine = shard_batch(batch, rank, world)
    m = SpatialVLAForConditionalGeneration(cfg, synth_state_dict(cfg, seed=0), ops=RefOps())
    m.engine.force_head = 0
    out = m.forward(**mine, attention_mask=torch.ones_like(mine["input_ids"]))
    res = reduce_loss(out)                     # the only collective of the labelled forward: 3 floats
    if rank == 0:
        q.put((res, float(out.loss)))
    dist.destroy_process_group()


def test_two_rank_loss_reduction_matches_single_process():
    """forward(labels) on a batch sharded over 2 ranks + parallel.reduce_loss == the single-process loss over the whole batch."""
    sys.path.insert(0, ROOT)
    import numpy as np
    from oracle.ops_ref import RefOps
    from spatialvla_b200.configs import get_config_dict
    from spatialvla_b200.modeling_spatialvla import SpatialVLAForConditionalGeneration
    from spatialvla_b200.parallel import reduce_loss
    from spatialvla_b200.weights import synth_state_dict
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_loss_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    (loss2, acc2, n2), loss_rank0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_model_train.npz"))
    cfg = get_config_dict("tiny")
    labels = torch.from_numpy(g["labels"])
    labels[1, -3:] = -100
    m = SpatialVLAForConditionalGeneration(cfg, synth_state_dict(cfg, seed=0), ops=RefOps())
    m.engine.force_head = 0
    ids = torch.from_numpy(g["input_ids"])
    out = m.forward(input_ids=ids, pixel_values=torch.from_numpy(g["pixel_u8"]).float() / 255.0, intrinsic=torch.from_numpy(g["intrinsic"]),
                    labels=labels, token_type_ids=torch.from_numpy(g["token_type_ids"]), attention_mask=torch.ones_like(ids))
    loss1, acc1, n1 = reduce_loss(out)         # no process group here: local values
    assert n2 == n1 == 11 and abs(loss1 - float(out.loss)) < 1e-5
    # bf16 activations: a batch of 1 and a batch of 2 take different fp32 GEMM blockings on the CPU, a few bf16 roundings flip
    assert abs(loss2 - loss1) < 5e-3 and abs(acc2 - acc1) < 1e-6, (loss2, loss1)
    assert abs(loss_rank0 - loss1) > 2e-2, (loss_rank0, loss1)      # rank 0's local mean is NOT the global one: the reduction did something
