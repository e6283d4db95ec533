"""N > 1 host logic on CPU with the gloo backend, world_size 2 (the data path has exactly one
collective: a SUM all-reduce of the flat LoRA gradient buffer, lora_b200/dist.py)."""
import os
import socket

import torch
import torch.multiprocessing as mp

from lora_b200.dist import shard_indices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from lora_b200.dist import allreduce_sum_, world_info
    from oracle import lora_ops as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert world_info() == (rank, world)
    torch.manual_seed(0)
    p0 = torch.randn(1000)                     # identical replica on every rank (same seed)
    torch.manual_seed(100 + rank)
    g_local = torch.randn(1000) * 3            # per-rank gradient from its own shard
    flat = g_local.clone()
    ws = allreduce_sum_(flat)
    assert ws == world
    # fused step semantics: g_sum with inv_world = 1/world  (clip on the AVERAGED gradient)
    new_p, _, _, total = O.clip_adamw_step([p0], [flat], [torch.zeros(1000)], [torch.zeros(1000)],
                                           1, [1e-3], inv_world=1.0 / world)
    torch.save({"p": new_p[0], "g_local": g_local, "total": total}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_then_step_equals_single_process_on_mean_gradient(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import lora_ops as O
    outs = [torch.load(tmp_path / f"r{r}.pt") for r in range(world)]
    assert torch.equal(outs[0]["p"], outs[1]["p"])            # replicas stay in lock-step
    g_mean = (outs[0]["g_local"] + outs[1]["g_local"]) / world
    torch.manual_seed(0)
    p0 = torch.randn(1000)
    want, _, _, total = O.clip_adamw_step([p0], [g_mean], [torch.zeros(1000)], [torch.zeros(1000)], 1, [1e-3])
    assert torch.allclose(outs[0]["p"], want[0], rtol=0, atol=1e-12)
    assert abs(outs[0]["total"] - total) < 1e-9


def test_shard_indices_round_robin():
    assert shard_indices(9, 0, 2) == [0, 2, 4, 6, 8]
    assert shard_indices(9, 1, 2) == [1, 3, 5, 7, 1]           # padded by wrapping: equal step counts
    assert shard_indices(9, 1, 2, drop_last=True) == [1, 3, 5, 7]
    assert shard_indices(4, 3, 8) == [3]
    for world in (1, 2, 4, 8):
        lens = {len(shard_indices(10, r, world)) for r in range(world)}
        assert len(lens) == 1
        seen = set()
        for r in range(world):
            seen.update(shard_indices(10, r, world))
        assert seen == set(range(10))
