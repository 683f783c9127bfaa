"""CPU, world_size 2 / 3 over gloo: the host logic of the rephrase-sharded decision -
shard boundaries, the score / action all-gather (even and ragged shards) and the rank-invariant selection.
The per-rank compute is replaced by a deterministic stand-in (no GPU here); the selection rule is the oracle's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cover_vla_b200.cover import CoverInputs, gather_and_select, rephrase_shard, shard_inputs
from oracle import verifier_oracle as V


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _select_cpu(scores, R, K):
    best, idx, gi, means = V.select(scores, K)
    return means, torch.tensor([idx], dtype=torch.int32), torch.tensor([best])


def _fake_inputs(R, K):
    g = torch.Generator().manual_seed(5)
    return CoverInputs(image=torch.zeros(3, 8, 8), lang_tokens=torch.arange(R * 4).view(R, 4), lang_len=torch.full((R,), 3, dtype=torch.int32),
                       state=torch.zeros(32), noise=torch.randn(R * K, 4, 32, generator=g), vf_image=torch.zeros(3, 8, 8),
                       vf_tokens=torch.zeros(4, dtype=torch.int64), past=None)


def _worker(rank, world, port, R, K, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = _fake_inputs(R, K)
        mine = shard_inputs(x, K, world, rank)
        a, b = rephrase_shard(R, world, rank)
        assert mine.lang_tokens.shape[0] == b - a and mine.noise.shape[0] == (b - a) * K
        assert torch.equal(mine.lang_tokens, x.lang_tokens[a:b])
        # stand-in for sample_and_score: a candidate's score / action depend only on its own noise row, so the
        # gathered result must not depend on how the candidates were sharded
        local_scores = mine.noise.flatten(1).sum(1).tanh()
        local_actions = mine.noise[:, :, :7].contiguous()
        scores, actions, gmean, idx, score = gather_and_select(local_scores, local_actions, R, K, _select_cpu)
        # numpy, not tensors: a tensor travels through the queue as a shared-memory handle that dies with this process -
        # if the worker exits before the parent has read it the parent's q.get() raises EOFError (seen under load)
        q.put((rank, scores.numpy().copy(), actions.numpy().copy(), int(idx), float(score)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,R,K", [(2, 8, 5), (2, 16, 16), (3, 8, 5), (2, 3, 2)])
def test_sharded_gather_and_select_is_rank_invariant(world, R, K):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = _fake_inputs(R, K)
    full_scores = x.noise.flatten(1).sum(1).tanh()
    best, idx, gi, means = V.select(full_scores, K)
    for rank, scores, actions, ridx, rscore in res:
        assert torch.equal(torch.from_numpy(scores), full_scores)
        assert torch.equal(torch.from_numpy(actions), x.noise[:, :, :7])
        assert ridx == idx and rscore == best


def test_rephrase_shards_partition_the_range():
    for R in (1, 3, 8, 16, 33):
        for G in (1, 2, 4, 8):
            spans = [rephrase_shard(R, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == R
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
