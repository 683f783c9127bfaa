"""Observation batching (BASELINE.json configs[4]): B independent decisions in one
cvb_pi0_sample_batch / cvb_cover_step_batch call must equal B single-observation calls - bit for bit on the same
(batch-capable) handle, whose kernels never let a row's result depend on how many observations share the launch - and
match the oracle like any single call."""
import pytest
import torch

from oracle import pi0_oracle as O
from oracle import verifier_oracle as V
from tests.helpers import action_gate, build_full_engine, build_pi0_engine, max_abs, pi0_truth

pytestmark = pytest.mark.gpu


def _obs(d, R, K, seed):
    inp = O.make_inputs(d, R, K, seed=seed)
    return inp, (inp["image"][0].cuda().contiguous(), inp["tokens"].cuda(), inp["lens"].to(torch.int32).cuda(),
                 inp["state"][0].cuda().contiguous(), inp["noise"].cuda())


@pytest.mark.parametrize("name,B,R,K", [("MID", 3, 2, 3), ("TINY", 4, 3, 2), ("MID", 8, 8, 5)])
def test_pi0_batch_equals_single_calls(name, B, R, K):
    d = getattr(O, name)
    w = O.make_pi0_weights(d, seed=3)
    eng = build_pi0_engine(d, w, R, K, max_observations=B)
    obs = [_obs(d, R, K, seed=20 + b) for b in range(B)]
    singles = [eng.pi0_sample(*a, K=K).clone() for _, a in obs]
    stack = [torch.stack([a[i] for _, a in obs]).contiguous() for i in range(5)]
    outs = [eng.pi0_sample_batch(*stack, K=K).clone() for _ in range(3)]  # eager, capture, replay
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    for b in range(B):
        assert torch.equal(outs[0][b], singles[b]), (b, max_abs(outs[0][b], singles[b]))
    # and the batched path is as right as the single-observation path: oracle at the reference batch layout + fp32 truth
    for b in range(min(B, 2)):
        inp = obs[b][0]
        bb = O.expand_to_batch(inp, K)
        ref = O.sample_actions(w, d, bb["image"], bb["tokens"], bb["masks"], bb["state"], bb["noise"])
        action_gate(outs[0][b].cpu(), ref, pi0_truth(O, w, d, inp, K), f"{name} batched observation {b}")
    eng.close()


def test_batch_handle_matches_latency_handle():
    """The same observation through a latency handle (split-K kernels, max_observations = 1) and a batch-capable handle
    (fused-epilogue GEMMs): same graph, different fp32 summation orders - both inside the action gate."""
    d, R, K = O.MID, 2, 3
    w = O.make_pi0_weights(d, seed=3)
    inp, args = _obs(d, R, K, seed=20)
    lat = build_pi0_engine(d, w, R, K)
    a = lat.pi0_sample(*args, K=K).cpu()
    lat.close()
    bat = build_pi0_engine(d, w, R, K, max_observations=4)
    b = bat.pi0_sample(*args, K=K).cpu()
    bat.close()
    bb = O.expand_to_batch(inp, K)
    ref = O.sample_actions(w, d, bb["image"], bb["tokens"], bb["masks"], bb["state"], bb["noise"])
    truth = pi0_truth(O, w, d, inp, K)
    action_gate(a, ref, truth, "latency handle")
    action_gate(b, ref, truth, "batch handle")
    assert max_abs(a, b) < 3e-2


@pytest.mark.parametrize("with_past", [False, True])
def test_cover_step_batch_equals_single_decisions(with_past):
    from cover_vla_b200.cover import BatchedCoverStep, CoverInputs, CoverStep
    d, v = O.MID, V.VMID
    B, R, K = 3, 2, 3
    w, vw = O.make_pi0_weights(d, 0), V.make_verifier_weights(v, 0)
    eng = build_full_engine(d, w, v, vw, R, K, max_observations=B)
    xs = []
    for b in range(B):
        inp = O.make_inputs(d, R, K, seed=40 + b)
        vin = V.make_inputs(v, 1, seed=40 + b)
        past = (torch.tensor([[0.01 * b, -0.02, 0.0, 0.03, 0.0, -0.05, 1.0]] * 2).cuda() if with_past else None)
        xs.append(CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                              lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                              noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(),
                              vf_tokens=vin["tokens"][0].cuda(), past=past))
    single = CoverStep(eng, K)
    refs = [[t.clone() for t in single.sample_and_score(x)] for x in xs]
    batched = BatchedCoverStep(eng, K)
    xb = BatchedCoverStep.stack(xs)
    for _ in range(3):  # eager, capture, replay
        out = [t.clone() for t in batched.sample_and_score(xb)]
        torch.cuda.synchronize()
        for b in range(B):
            for got, ref in zip(out, refs[b]):
                assert torch.equal(got[b].reshape(-1), ref.reshape(-1)), (b, max_abs(got[b].reshape(-1), ref.reshape(-1)))
    idx, score, win = batched(xb, gate_threshold=10.0)
    for b in range(B):
        i1, s1, w1 = single(xs[b], gate_threshold=10.0)
        assert int(idx[b]) == i1 and float(score[b]) == s1 and torch.equal(win[b], w1)
    eng.close()
