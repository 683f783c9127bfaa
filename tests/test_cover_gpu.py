"""Whole CoVer decision on the GPU through the public host surfaces vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import pi0_oracle as O
from oracle import verifier_oracle as V
from tests.helpers import SCORE_TOL, action_gate, build_full_engine, max_abs, pi0_truth, score_gate, verifier_truth_scores

pytestmark = pytest.mark.gpu


def _oracle_traj(actions, past, history, p01, p99):
    """process_inputs(verifier_action=True) + postprocess_verifier + padding: the oracle restatement, pinned by
    tests/golden/format_traj.npz (tests/test_format_traj.py)."""
    from oracle import exec_action_oracle as X
    return torch.from_numpy(X.verifier_trajectories(actions.cpu().numpy(), None if past is None else past.cpu().numpy(),
                                                    history, p01, p99))


def test_cover_step_matches_oracle():
    from cover_vla_b200.cover import BRIDGE_ACTION_P01, BRIDGE_ACTION_P99, CoverInputs, CoverStep
    d, v = O.TINY, V.VTINY
    R, K = 4, 3
    w, vw = O.make_pi0_weights(d, 0), V.make_verifier_weights(v, 0)
    eng = build_full_engine(d, w, v, vw, R, K)
    inp = O.make_inputs(d, R, K, seed=5)
    vin = V.make_inputs(v, 1, seed=5)
    past = torch.tensor([[0.01, -0.02, 0.0, 0.03, 0.0, -0.05, 1.0], [0.0, 0.01, 0.02, 0.0, 0.0, 0.1, 0.0]])
    x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                    lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                    noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(),
                    vf_tokens=vin["tokens"][0].cuda(), past=past.cuda())
    step = CoverStep(eng, K)
    actions, traj, scores, gmean, bidx, bscore = step.sample_and_score(x)
    torch.cuda.synchronize()
    # formatting kernel == numpy restatement, bit for bit (given the same actions)
    ref_traj = _oracle_traj(actions, past, v.history, BRIDGE_ACTION_P01, BRIDGE_ACTION_P99)
    assert torch.equal(traj.cpu(), ref_traj)
    # verifier on those trajectories
    patch, text = V.extract_features(vw, v, vin["image"], vin["tokens"])
    ref_scores = V.scores_from_features(vw, v, patch, text, ref_traj)
    best, idx, gi, means = V.select(ref_scores, K)
    err = score_gate(scores, ref_scores, verifier_truth_scores(V, vw, v, vin["image"], vin["tokens"], ref_traj), "cover step")
    print(f"cover step: best idx {int(bidx.item())} (oracle {idx})")
    tol = max(SCORE_TOL, 2 * err)
    # gate (run_simpler_eval_with_openpi.py:344-363): candidate 0 is kept iff its score >= threshold
    i2, s2, win = step(x, gate_threshold=10.0)   # score < 10 -> the N-candidate selection is used
    assert i2 == idx and win.shape == (d.chunk_size, 7) and abs(s2 - best) <= tol
    assert torch.equal(win, actions[idx, :, :7].cpu())
    i3, s3, win3 = step(x, gate_threshold=-1e9)  # always confident -> candidate 0
    assert i3 == 0 and abs(s3 - float(ref_scores[0])) <= tol
    # the action the reference executes (run_simpler_eval_with_openpi.py:368-391): execution format + gripper vote
    from oracle import exec_action_oracle as X
    i4, s4, win4, ex = step.decide_and_execute(x, gate_threshold=10.0)
    assert i4 == idx and torch.equal(win4, win)
    ex_ref, _ = X.execution_action(actions.cpu().numpy(), idx, K, BRIDGE_ACTION_P01, BRIDGE_ACTION_P99)
    assert np.array_equal(ex[:3], ex_ref[:3]) and ex[6] == ex_ref[6]
    assert np.allclose(ex[3:6], ex_ref[3:6], rtol=0, atol=1e-12)
    eng.close()


def test_prompt_cache_hold_text_is_exact():
    """cvb_verifier_hold_text (per-task prompt cache): with the instruction unchanged the decision
    with the text tower skipped equals the full one bit for bit - fused step, eager and graph replay, single and batched;
    releasing the hold re-encodes a new instruction."""
    from cover_vla_b200.cover import BatchedCoverStep, CoverInputs, CoverStep
    d, v = O.TINY, V.VTINY
    R, K, B = 2, 2, 2
    eng = build_full_engine(d, O.make_pi0_weights(d, 0), v, V.make_verifier_weights(v, 0), R, K, max_observations=B)
    xs = []
    for b in range(B):
        inp = O.make_inputs(d, R, K, seed=20 + b)
        vin = V.make_inputs(v, 1, seed=20 + b)
        xs.append(CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                              lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                              noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(),
                              vf_tokens=vin["tokens"][0].cuda(), past=None))
    step = CoverStep(eng, K)
    base = [t.clone() for t in step.sample_and_score(xs[0])]
    step.hold_text = True
    for _ in range(3):  # eager, capture, replay of the held variant
        held = step.sample_and_score(xs[0])
        assert all(torch.equal(a, b) for a, b in zip(held, base))
    # a new image with the same instruction: still exact against the un-held computation
    x2 = CoverInputs(**{**xs[0].__dict__, "vf_image": xs[1].vf_image})
    held2 = [t.clone() for t in step.sample_and_score(x2)]
    step.hold_text = False
    full2 = step.sample_and_score(x2)
    assert all(torch.equal(a, b) for a, b in zip(held2, full2))
    assert not torch.equal(full2[2], base[2])
    # batched: B observations
    bstep = BatchedCoverStep(eng, K)
    xb = BatchedCoverStep.stack(xs)
    bbase = [t.clone() for t in bstep.sample_and_score(xb)]
    bstep.hold_text = True
    for _ in range(3):
        assert all(torch.equal(a, b) for a, b in zip(bstep.sample_and_score(xb), bbase))
    # a single-observation call in between (graph replay) overwrites part of the resident text features: the engine
    # tracks the observation count of the last context on the host path and re-encodes for the batch although held
    step.hold_text = False
    step.sample_and_score(xs[1])
    assert all(torch.equal(a, b) for a, b in zip(bstep.sample_and_score(xb), bbase))
    eng.close()


def test_policy_with_two_cameras_and_a_missing_one():
    """PI0Policy.select_action with two image keys (prepare_images, modeling_pi0.py:344-387): both present -> two image
    streams; one key missing from the batch -> the reference would add a masked -1 image (a no-op), the mirror runs the
    present camera alone."""
    from cover_vla_b200.pi0 import PI0Config, PI0Policy, PolicyFeature
    d = O.TINY
    R, K = 2, 2
    N = R * K
    w = O.make_pi0_weights(d, 0)
    feat = PolicyFeature("VISUAL", (3, d.vis_image, d.vis_image))
    cfg = PI0Config(chunk_size=d.chunk_size, n_action_steps=d.chunk_size, tokenizer_max_length=d.max_lang_len,
                    proj_width=d.ex_width, num_steps=d.num_steps, vis_layers=d.vis_layers, vis_width=d.vis_width,
                    vis_heads=d.vis_heads, vis_mlp=d.vis_mlp, vis_patch=d.vis_patch, vis_image=d.vis_image,
                    layers=d.layers, lm_width=d.lm_width, lm_mlp=d.lm_mlp, heads=d.heads, head_dim=d.head_dim,
                    ex_mlp=d.ex_mlp, vocab=d.vocab, max_rephrases=R, max_samples=K, empty_cameras=1,
                    resize_imgs_with_padding=(d.vis_image, d.vis_image),
                    input_features={"observation.images.top": feat, "observation.images.wrist": feat,
                                    "observation.state": PolicyFeature("STATE", (7,))})
    policy = PI0Policy(cfg, state_dict={"model." + k: t for k, t in w.items()})
    assert policy.model.engine.cfg.num_cameras == 2
    inp = O.make_inputs(d, R, K, seed=9)
    b = O.expand_to_batch(inp, K)
    cam2 = (torch.rand(1, 3, d.vis_image, d.vis_image, generator=torch.Generator().manual_seed(78)) * 2 - 1).repeat(N, 1, 1, 1)
    obs = {"observation.images.top": b["image"].cuda(), "observation.images.wrist": cam2.cuda(),
           "observation.state": b["state"][:, :7].cuda(), "lang_tokens": b["tokens"].cuda(),
           "lang_masks": b["masks"].cuda(), "task": ["x"] * N}
    q = policy.select_action(obs, noise=b["noise"].cuda())
    got = torch.stack(list(q), dim=1).cpu()
    q.clear()
    ref = O.sample_actions(w, d, [b["image"], cam2], b["tokens"], b["masks"], b["state"], b["noise"])
    with O.truth_mode():
        truth = O.sample_actions_dedup(O.truth_weights(w), d, [inp["image"], cam2[:1]], inp["tokens"], inp["masks"],
                                       inp["state"], inp["noise"], K)
    action_gate(got, ref[:, :, :7], truth[:, :, :7], "select_action, 2 cameras")
    obs.pop("observation.images.wrist")  # missing key = empty camera
    q = policy.select_action(obs, noise=b["noise"].cuda())
    got1 = torch.stack(list(q), dim=1).cpu()
    # = the engine with one active camera on the same handle, bit for bit (the numerical parity of that call against the
    # oracle is tests/test_pi0_gpu.py::test_two_cameras_match_oracle_and_masked_camera_is_dropped; that the reference's
    # masked -1 image is a no-op is tests/test_oracle_vs_reference.py::test_pi0_two_cameras_and_masked_camera_vs_reference)
    eng = policy.model.engine
    assert getattr(eng, "_active_cams", 0) == 1
    direct = eng.pi0_sample(inp["image"][0].cuda().contiguous(), inp["tokens"].cuda(), inp["lens"].to(torch.int32).cuda(),
                            torch.nn.functional.pad(inp["state"][0, :7], (0, d.max_state_dim - 7)).cuda().contiguous(),
                            inp["noise"].cuda(), K=K).cpu()
    assert torch.equal(got1, direct[:, :, :7])
    one = O.sample_actions(w, d, b["image"], b["tokens"], b["masks"], b["state"], b["noise"])
    assert max_abs(got1, one[:, :, :7]) < 3e-2 < max_abs(got1, ref[:, :, :7])  # it is the 1-camera result, not the 2-camera one


def test_policy_and_ensemble_surfaces():
    """The two reference-facing objects, used the way run_simpler_eval_with_openpi.py uses them."""
    from cover_vla_b200.pi0 import PI0Config, PI0Policy, PolicyFeature
    from cover_vla_b200.verifier import EfficientEnsembleMerged
    d, v = O.TINY, V.VTINY
    R, K = 3, 2
    N = R * K
    w, vw = O.make_pi0_weights(d, 0), V.make_verifier_weights(v, 0)
    cfg = PI0Config(chunk_size=d.chunk_size, n_action_steps=d.chunk_size, tokenizer_max_length=d.max_lang_len,
                    proj_width=d.ex_width, num_steps=d.num_steps, vis_layers=d.vis_layers, vis_width=d.vis_width,
                    vis_heads=d.vis_heads, vis_mlp=d.vis_mlp, vis_patch=d.vis_patch, vis_image=d.vis_image,
                    layers=d.layers, lm_width=d.lm_width, lm_mlp=d.lm_mlp, heads=d.heads, head_dim=d.head_dim,
                    ex_mlp=d.ex_mlp, vocab=d.vocab, max_rephrases=R, max_samples=K,
                    resize_imgs_with_padding=(d.vis_image, d.vis_image),
                    input_features={"observation.images.top": PolicyFeature("VISUAL", (3, d.vis_image, d.vis_image)),
                                    "observation.state": PolicyFeature("STATE", (7,))})
    policy = PI0Policy(cfg, state_dict={"model." + k: t for k, t in w.items()})
    inp = O.make_inputs(d, R, K, seed=7)
    b = O.expand_to_batch(inp, K)
    obs = {"observation.images.top": b["image"].cuda(), "observation.state": b["state"][:, :7].cuda(),
           "lang_tokens": b["tokens"].cuda(), "lang_masks": b["masks"].cuda(), "task": ["x"] * N}
    q = policy.select_action(obs, noise=b["noise"].cuda())
    assert len(q) == cfg.n_action_steps and q[0].shape == (N, 7)
    queue = q.copy()
    q.clear()
    ref = O.sample_actions(w, d, b["image"], b["tokens"], b["masks"], b["state"], b["noise"])
    got = torch.stack(list(queue), dim=1).cpu()  # [N, steps, 7]
    action_gate(got, ref[:, :, :7], pi0_truth(O, w, d, inp, K)[:, :, :7], "select_action TINY")
    # a second call with an empty queue samples again; with a non-empty queue it must not
    q2 = policy.select_action(obs, noise=b["noise"].cuda())
    assert len(q2) == cfg.n_action_steps
    q3 = policy.select_action(obs, noise=torch.zeros_like(b["noise"]).cuda())
    assert torch.equal(torch.stack(list(q3)), torch.stack(list(q2)))

    comps = []
    for m in range(v.members):
        c = {}
        for k, t in vw.items():
            pre = f"verifier.{m}."
            if k.startswith(pre):
                comp, name = k[len(pre):].split(".", 1)
                c.setdefault(comp, {})[name] = t
        c["action_padding_value"] = -5.0
        comps.append(c)
    trunk = {k: t for k, t in vw.items() if k.startswith("verifier.trunk.")}
    vf_cfg = dict(vf_image=v.image, vf_patch=v.patch, vf_width=v.width, vf_layers=v.layers, vf_heads=v.heads,
                  vf_mlp=v.mlp, vf_text_layers=v.text_layers, vf_text_ctx=v.text_ctx, vf_vocab=v.vocab,
                  vf_embed=v.embed, vf_pool_heads=v.pool_heads, vf_pool_layers=v.pool_layers,
                  vf_traj_layers=v.traj_layers, vf_traj_ff=v.traj_ff)
    ens = EfficientEnsembleMerged(ensemble_components=comps, trunk_state_dict=trunk, vf_config=vf_cfg,
                                  preprocess=lambda im: im, max_candidates=N)
    vin = V.make_inputs(v, N, seed=7)
    imgs = [vin["image"][0]] * N
    instr = [vin["tokens"][0]] * N
    ms, mi, mh, gi = ens.compute_max_similarity_scores_batch(imgs, instr, vin["histories"], cfg_repeat_language_instructions=K)
    best, idx, ref_scores, means = V.compute_max_similarity_scores(vw, v, vin["image"], vin["tokens"], vin["histories"], K)
    assert isinstance(ms, float) and gi.dtype == torch.int64 and gi.ndim == 0
    hist, sc = ens.predict(vin["image"][0], vin["tokens"][0], vin["histories"])
    got_scores = torch.tensor([sc[str(i)] for i in range(N)])
    traj_ref = V.pad_histories(vin["histories"], v.history)
    err = score_gate(got_scores, ref_scores, verifier_truth_scores(V, vw, v, vin["image"], vin["tokens"], traj_ref), "ensemble surface")
    tol = max(SCORE_TOL, 2 * err)
    assert abs(ms - best) <= tol
    assert mh is vin["histories"][int(gi)]
    # the 1-candidate gate call reuses the context and equals scores[0]
    ms1, _, _, gi1 = ens.compute_max_similarity_scores_batch(imgs[:1], instr[:1], vin["histories"][:1], cfg_repeat_language_instructions=1)
    assert int(gi1) == 0 and abs(ms1 - float(ref_scores[0])) <= tol
    assert len(sc) == N and abs(sc["0"] - float(ref_scores[0])) <= tol
    # a uint8 frame through the DEFAULT transform takes the device-side open_clip transform (bit-exact with PIL), so the
    # scores equal those of the host-side PIL route exactly
    from PIL import Image
    from cover_vla_b200.verifier.efficient_ensemble_merged import default_preprocess
    frame = np.random.default_rng(0).integers(0, 256, size=(96, 128, 3)).astype(np.uint8)
    ens2 = EfficientEnsembleMerged(ensemble_components=comps, trunk_state_dict=trunk, vf_config=vf_cfg, max_candidates=N)
    s_dev = ens2.compute_max_similarity_scores_batch([Image.fromarray(frame)] * N, instr, vin["histories"],
                                                     cfg_repeat_language_instructions=K)[0]
    host_img = default_preprocess(v.image)(Image.fromarray(frame))
    s_host = ens.compute_max_similarity_scores_batch([host_img] * N, instr, vin["histories"],
                                                     cfg_repeat_language_instructions=K)[0]
    assert s_dev == s_host
    policy.engine.close()
    ens.engine.close()
    ens2.engine.close()


@pytest.mark.parametrize("with_past", [True, False])
def test_fused_cover_step_equals_three_calls(with_past):
    """cvb_cover_step (one graph, verifier context forked after the prefix) == cvb_pi0_sample + cvb_format_trajectories
    + cvb_verifier_score, bit for bit, eagerly and on graph replay."""
    from cover_vla_b200.cover import CoverInputs, CoverStep
    d, v = O.MID, V.VMID
    R, K = 3, 2
    w, vw = O.make_pi0_weights(d, 0), V.make_verifier_weights(v, 0)
    eng = build_full_engine(d, w, v, vw, R, K)
    inp = O.make_inputs(d, R, K, seed=9)
    vin = V.make_inputs(v, 1, seed=9)
    past = torch.tensor([[0.01, -0.02, 0.0, 0.03, 0.0, -0.05, 1.0]] * 3).cuda() if with_past else None
    x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                    lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                    noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(),
                    vf_tokens=vin["tokens"][0].cuda(), past=past, lang_len_max=int(inp["lens"].max()))
    step = CoverStep(eng, K)
    step.fused = False
    ref = [t.clone() for t in step.sample_and_score(x)]
    step.fused = True
    for _ in range(4):  # eager, capture, replay, replay
        out = step.sample_and_score(x)
        torch.cuda.synchronize()
        for a, b in zip(out, ref):
            assert torch.equal(a, b)
    eng.close()


@pytest.mark.parametrize("R,K", [(3, 2), (4, 3)])
def test_rephrase_shards_equal_the_whole_decision(R, K):
    """The sharding invariant of SURVEY.md section 8e on ONE GPU: what a rank computes for its rephrase slice - also a
    slice of ONE rephrase - is bit-identical to those candidates' rows in the whole decision, for every world size.  (No
    kernel choice may depend on how many prompts share a launch: round 2 briefly picked the RMSNorm kernel by row count,
    which only the 2-GPU test tests/test_sharded_nccl_gpu.py - skipped on a 1-GPU box - could see.)"""
    from cover_vla_b200.cover import CoverInputs, CoverStep, shard_inputs
    d, v = O.MID, V.VMID
    w, vw = O.make_pi0_weights(d, 0), V.make_verifier_weights(v, 0)
    eng = build_full_engine(d, w, v, vw, R, K)
    inp = O.make_inputs(d, R, K, seed=21)
    vin = V.make_inputs(v, 1, seed=21)
    x = CoverInputs(image=inp["image"][0].cuda().contiguous(), lang_tokens=inp["tokens"].cuda(),
                    lang_len=inp["lens"].to(torch.int32).cuda(), state=inp["state"][0].cuda().contiguous(),
                    noise=inp["noise"].cuda(), vf_image=vin["image"][0].cuda().contiguous(),
                    vf_tokens=vin["tokens"][0].cuda(), past=None, lang_len_max=int(inp["lens"].max()))
    step = CoverStep(eng, K)
    whole = [t.clone() for t in step.sample_and_score(x)]
    for world in range(2, R + 1):
        parts = [[t.clone() for t in step.sample_and_score(shard_inputs(x, K, world, r))] for r in range(world)]
        for i, name in enumerate(["actions", "trajectories", "scores"]):
            got = torch.cat([p[i] for p in parts])
            assert torch.equal(got, whole[i]), (world, name, max_abs(got, whole[i]))
    eng.close()
