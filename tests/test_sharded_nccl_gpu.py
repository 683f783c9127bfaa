"""Two ranks over NCCL (one process per GPU): the rephrase-sharded decision returns, on every rank, exactly what one
GPU returns for the whole candidate set.  Needs >= 2 CUDA devices (skipped on a 1-GPU box; run with gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, R, K, peer, q):
    import torch.distributed as dist
    from cover_vla_b200 import synthetic as S
    from cover_vla_b200.cover import CoverInputs, CoverStep, ShardedCoverStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        d, v = S.MID, S.VMID
        w, vw = S.make_pi0_weights(d, 0), S.make_verifier_weights(v, 0)
        dev = f"cuda:{rank}"
        eng = S.build_engine(d, w, v, vw, R, K, device=dev)
        inp = S.make_inputs(d, R, K, seed=21)
        vin = S.make_verifier_inputs(v, 1, seed=21)
        x = CoverInputs(image=inp["image"][0].to(dev).contiguous(), lang_tokens=inp["tokens"].to(dev),
                        lang_len=inp["lens"].to(torch.int32).to(dev), state=inp["state"][0].to(dev).contiguous(),
                        noise=inp["noise"].to(dev), vf_image=vin["image"][0].to(dev).contiguous(),
                        vf_tokens=vin["tokens"][0].to(dev), past=None, lang_len_max=int(inp["lens"].max()))
        sstep = ShardedCoverStep(eng, K, peer_memory=peer)
        assert (sstep.peer is not None) == peer
        for _ in range(4):  # several decisions: the mailbox epochs / parities advance, results must not
            scores, actions, gmean, idx, score = sstep(x)
            torch.cuda.synchronize()
        # numpy, not tensors: a CPU tensor travels through the queue as a shared-memory handle that dies with this process
        out = dict(rank=rank, scores=scores.cpu().numpy(), actions=actions.cpu().numpy(), idx=int(idx.item()),
                   score=float(score.item()))
        if rank == 0:  # single-GPU answer for the whole candidate set
            a1, t1, s1, g1, i1, b1 = CoverStep(eng, K).sample_and_score(x)
            torch.cuda.synchronize()
            out.update(ref_scores=s1.cpu().numpy(), ref_actions=a1[:, :, :7].cpu().numpy(), ref_idx=int(i1.item()),
                       ref_score=float(b1.item()))
        q.put(out)
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("R,K,peer", [(4, 3, True), (3, 2, True), (4, 3, False)])
def test_sharded_decision_equals_single_gpu(R, K, peer):
    """peer=True: cvb_allgather_select (peer-memory stores over NVLink + in-kernel selection); False: NCCL all-gathers."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, K, peer, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda o: o["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = res[0]
    for o in res:
        # per-candidate results do not depend on which rank computed them: bit-identical to the one-GPU run
        assert np.array_equal(o["scores"], ref["ref_scores"])
        assert np.array_equal(o["actions"], ref["ref_actions"])
        assert o["idx"] == ref["ref_idx"] and o["score"] == ref["ref_score"]
