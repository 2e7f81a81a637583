"""CPU-only, world_size 2 over gloo: the multi-rank path of bench.py / the sharded dataset run.
Each rank scores its own round-robin shard of blocks (no data-path collective); gathering the
shards back must reproduce the single-process result row for row, and timing is max over ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from anyscale_workshop_nyc_2023_b200.parallel import max_over_ranks, restore_order, shard_block_indices
from anyscale_workshop_nyc_2023_b200.synth import SPECS, make_state_dict, synthetic_token_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _score(block_ids, block_mask):
    from oracle.t5_oracle import T5Oracle

    o = T5Oracle(make_state_dict(SPECS["tiny"], 1), SPECS["tiny"], emulate_bf16=False)
    return o.generate(block_ids, block_mask, max_new_tokens=4, min_new_tokens=4)[0]


def _worker(rank, world, port, n_blocks, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    ids, mask = synthetic_token_batch(n_blocks * 2, 8, SPECS["tiny"].vocab_size, seed=4, lengths="uniform")
    mine = shard_block_indices(n_blocks, rank, world)
    outs = [_score(ids[2 * i: 2 * i + 2], mask[2 * i: 2 * i + 2]) for i in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, outs)  # test-only: the product path never gathers on the data path
    t = max_over_ranks(float(rank + 1))
    if rank == 0:
        q.put((np.concatenate(restore_order(gathered, n_blocks)), t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sharded_run_equals_single_process():
    n_blocks, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_blocks, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids, mask = synthetic_token_batch(n_blocks * 2, 8, SPECS["tiny"].vocab_size, seed=4, lengths="uniform")
    want = np.concatenate([_score(ids[2 * i: 2 * i + 2], mask[2 * i: 2 * i + 2]) for i in range(n_blocks)])
    assert (got == want).all()
    assert tmax == 2.0
