"""Parity at the BASELINE.json configurations of record - FULL-size pi0 (3.5 B parameters) and the full-size verifier -
against fixtures produced by the UNMODIFIED reference files in the authoring container (oracle/make_golden.py,
oracle/make_golden_verifier.py; the GPU box has no /root/reference and the CPU oracle needs minutes at this size).

  configs[1]  pi0 sample_actions, 5 samples x 1 instruction            tests/golden/pi0_full_R1K5.pt
  configs[2]  pi0 8 rephrases x 5 samples (40 candidates)              tests/golden/pi0_full_R8K5.pt (+ R2K2)
  configs[0]  bridge_verifier scoring, 8 x 5 trajectories, full size   tests/golden/verifier_vfull_R8K5.pt

Each pi0 fixture also carries the fp32 'truth' of the same graph (oracle truth_mode), so the action
gate is the stated 1e-2 against the reference OR, where the reference's own bf16 error against the truth is larger
than that, 'no further from the truth than 1.2 x the reference' - all three numbers are printed.

One engine (pi0 + verifier) is shared by the module: generating 3.5 B synthetic parameters takes about a minute.
"""
from pathlib import Path

import pytest
import torch

from oracle import pi0_oracle as O
from oracle import verifier_oracle as V
from tests.helpers import SCORE_TOL, action_gate, build_full_engine, max_abs, rel_l2

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def full_engine():
    d, v = O.FULL, V.VFULL
    w = O.make_pi0_weights(d, 0)
    vw = V.make_verifier_weights(v, 0)
    eng = build_full_engine(d, w, v, vw, 8, 5)
    del w
    yield eng, vw
    eng.close()


def _pi0_args(inp):
    return (inp["image"][0].cuda().contiguous(), inp["tokens"].cuda(), inp["lens"].to(torch.int32).cuda(),
            inp["state"][0].cuda().contiguous(), inp["noise"].cuda())


@pytest.mark.parametrize("R,K", [(2, 2), (1, 5), (8, 5)])
def test_pi0_full_size_against_reference_golden(full_engine, R, K):
    eng, _ = full_engine
    d = O.FULL
    g = torch.load(GOLD / f"pi0_full_R{R}K{K}.pt")
    assert g["dims"] == d.as_dict() and g["R"] == R and g["K"] == K
    inp = O.make_inputs(d, R, K, seed=g["seed"])
    assert torch.equal(inp["lens"], g["lens"])
    outs = [eng.pi0_sample(*_pi0_args(inp), K=K).cpu() for _ in range(3)]  # eager, capture, replay
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    T, P, L = d.n_img_tokens, d.n_img_tokens + d.max_lang_len, d.layers - 1
    # stage checks on the strided slices the fixture keeps (relative L2: bf16 rounding noise)
    img = eng.debug("image_emb", (T, d.lm_width), torch.bfloat16).float() / (d.lm_width ** 0.5)
    e_img = rel_l2(img[::37, ::29], g["image_emb_slice"])
    k0 = eng.debug("prefix_k0", (8, P, d.head_dim), torch.bfloat16)[:R]
    vl = eng.debug("prefix_vlast", (8, P, d.head_dim), torch.bfloat16)[:R]
    e_k0 = e_vl = 0.0
    for r in range(R):  # valid tokens only (the engine never computes the right-padding rows)
        n = T + int(inp["lens"][r])
        rows = torch.arange(0, P, 11)
        rows = rows[rows < n]
        e_k0 = max(e_k0, rel_l2(k0[r, rows, ::7], g["k0_slice"][r, : len(rows)]))
        e_vl = max(e_vl, rel_l2(vl[r, rows, ::7], g["vlast_slice"][r, : len(rows)]))
    v0 = eng.debug("v0", (8 * 5, d.chunk_size, d.max_action_dim), torch.float32)[: R * K]
    e_v0 = rel_l2(v0, g["v0"])
    print(f"FULL R={R} K={K}: rel-L2 image_emb {e_img:.2e}  k0 {e_k0:.2e}  v_last {e_vl:.2e}  v0 {e_v0:.2e}")
    assert e_img < 2e-2 and e_k0 < 2e-2 and e_vl < 5e-2 and e_v0 < 5e-2
    action_gate(outs[0], g["actions"], g["actions_truth"], f"FULL R={R} K={K} actions")
    assert abs(g["err_ref_vs_truth"] - max_abs(g["actions"], g["actions_truth"])) < 1e-6


def test_verifier_full_size_against_reference_golden(full_engine):
    eng, vw = full_engine
    v = V.VFULL
    R, K = 8, 5
    g = torch.load(GOLD / "verifier_vfull_R8K5.pt")
    inp = V.make_inputs(v, R * K, seed=g["seed"])
    traj = V.pad_histories(inp["histories"], v.history)
    scores, gmean, bidx, bscore = eng.verifier_score(inp["image"][0].cuda().contiguous(), inp["tokens"][0].cuda(),
                                                     traj.cuda(), R, K)
    torch.cuda.synchronize()
    pf = eng.debug("vf_patch_features", (v.n_patches, v.width), torch.float32)
    tf = eng.debug("vf_text_features", (v.text_ctx, v.width), torch.float32)
    e_p, e_t = rel_l2(pf[::7, ::13], g["patch_slice"]), rel_l2(tf[::3, ::13], g["text_slice"])
    it = eng.debug("vf_it_emb", (v.members, v.embed), torch.float32)
    e_it = rel_l2(it, g["it_emb"])
    err = max_abs(scores, g["scores"])
    ref = g["scores"]
    srt = torch.sort(ref.view(R, K)[g["global_idx"] // K], descending=True).values
    means = torch.sort(ref.view(R, K).mean(1), descending=True).values
    gap, ggap = (srt[0] - srt[1]).item(), (means[0] - means[1]).item()
    print(f"VFULL R8K5: trunk rel-L2 patch {e_p:.2e} text {e_t:.2e}, image-text embedding rel-L2 {e_it:.2e}, scores max-abs "
          f"{err:.2e} (|score|max {ref.abs().max().item():.2e}), idx {int(bidx.item())} vs reference {g['global_idx']}, "
          f"gaps {gap:.2e} / {ggap:.2e}")
    assert e_p < 3e-2 and e_t < 3e-2 and e_it < 1e-2
    assert err < SCORE_TOL
    if gap > 2 * err and ggap > 2 * err:
        assert int(bidx.item()) == g["global_idx"]
        assert abs(bscore.item() - g["max_score"]) < SCORE_TOL


def test_cover_step_full_size_is_consistent(full_engine):
    """configs[2] end to end: the fused decision (one graph, verifier context forked beside the sampler) returns exactly
    the actions of cvb_pi0_sample and exactly the scores of cvb_verifier_score on its own trajectories."""
    from cover_vla_b200.cover import CoverInputs, CoverStep
    eng, _ = full_engine
    d, v = O.FULL, V.VFULL
    R, K = 8, 5
    g = torch.load(GOLD / "pi0_full_R8K5.pt")
    inp = O.make_inputs(d, R, K, seed=g["seed"])
    vin = V.make_inputs(v, 1, seed=2)
    x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                    lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                    noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(),
                    vf_tokens=vin["tokens"][0].cuda(), past=None)
    step = CoverStep(eng, K)
    actions, traj, scores, gmean, bidx, bscore = [t.clone() for t in step.sample_and_score(x)]
    torch.cuda.synchronize()
    action_gate(actions.cpu(), g["actions"], g["actions_truth"], "FULL cover step actions")
    s2, gm2, bi2, bs2 = eng.verifier_score(x.vf_image, x.vf_tokens, traj, R, K)
    assert torch.equal(scores, s2) and int(bidx.item()) == int(bi2.item())
    sel = scores.view(R, K)
    gi = int(sel.mean(1).argmax())
    assert int(bidx.item()) == gi * K + int(sel[gi].argmax())
