"""CPU, authoring container only: the oracle restatements against the UNMODIFIED reference files executed live
through oracle/ref_shim.py.  Skipped wherever /root/reference is absent (the GPU box); the committed fixtures in
tests/golden/ (tests/test_oracle_golden.py) carry the same pin there."""
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present")


def _ref_model(d, seed):
    from oracle import pi0_oracle as O
    torch.manual_seed(0)
    model, _ = ref_shim.build_pi0(d.as_dict(), chunk_size=d.chunk_size, tokenizer_max_length=d.max_lang_len,
                                  num_steps=d.num_steps)
    w = O.make_pi0_weights(d, seed=seed)
    sd = model.state_dict()
    for k, v in w.items():
        sd[O.to_hf5_key(k)].copy_(v)
    return model, w


@pytest.mark.parametrize("R,K,seed", [(2, 2, 0), (3, 1, 4), (1, 3, 5)])
def test_pi0_sample_actions_bit_exact_vs_reference(R, K, seed):
    """PI0FlowMatching.sample_actions, modeling_pi0.py:672-715, tiny config, reference batch layout."""
    from oracle import pi0_oracle as O
    d = O.TINY
    model, w = _ref_model(d, seed)
    inp = O.make_inputs(d, R, K, seed=seed)
    b = O.expand_to_batch(inp, K)
    N = R * K
    with torch.no_grad():
        ref = model.sample_actions([b["image"]], [torch.ones(N, dtype=torch.bool)], b["tokens"], b["masks"], b["state"],
                                   noise=b["noise"].clone())
    out = O.sample_actions(w, d, b["image"], b["tokens"], b["masks"], b["state"], b["noise"])
    assert torch.equal(out, ref)


def test_pi0_two_cameras_and_masked_camera_vs_reference():
    """prepare_images / embed_prefix with several image streams (modeling_pi0.py:344-387, 529-547): the oracle with a list
    of cameras is bit-exact against the reference, and a masked ("empty") camera is a no-op - the reference with the -1
    image and an all-False mask equals the reference without that camera (what the CUDA path relies on to drop it)."""
    from oracle import pi0_oracle as O
    d = O.TINY
    R, K, seed = 2, 2, 3
    model, w = _ref_model(d, seed)
    inp = O.make_inputs(d, R, K, seed=seed)
    b = O.expand_to_batch(inp, K)
    N = R * K
    cam2 = torch.rand(1, 3, d.vis_image, d.vis_image, generator=torch.Generator().manual_seed(77)) * 2 - 1
    cam2 = cam2.repeat(N, 1, 1, 1)
    on, off = torch.ones(N, dtype=torch.bool), torch.zeros(N, dtype=torch.bool)
    with torch.no_grad():
        ref2 = model.sample_actions([b["image"], cam2], [on, on], b["tokens"], b["masks"], b["state"], noise=b["noise"].clone())
        ref1 = model.sample_actions([b["image"]], [on], b["tokens"], b["masks"], b["state"], noise=b["noise"].clone())
        refm = model.sample_actions([b["image"], torch.ones_like(cam2) * -1], [on, off], b["tokens"], b["masks"], b["state"],
                                    noise=b["noise"].clone())
    out2 = O.sample_actions(w, d, [b["image"], cam2], b["tokens"], b["masks"], b["state"], b["noise"], img_masks=[on, on])
    assert torch.equal(out2, ref2)
    outm = O.sample_actions(w, d, [b["image"], torch.ones_like(cam2) * -1], b["tokens"], b["masks"], b["state"], b["noise"],
                            img_masks=[on, off])
    assert torch.equal(outm, refm)
    assert (ref2 - ref1).abs().max().item() > 1e-2      # the second camera matters ...
    assert (refm - ref1).abs().max().item() <= 1e-2     # ... a masked one does not (bf16 GEMM blocking noise only)


def test_pi0_dedup_is_exact_on_the_reference_itself():
    """one prefix per unique rephrase shared by its K samples, padded language tokens dropped -
    the de-duplicated schedule the CUDA path runs equals the reference's B = N batch (bit-exact on CPU)."""
    from oracle import pi0_oracle as O
    d = O.TINY
    R, K = 2, 3
    model, w = _ref_model(d, 2)
    inp = O.make_inputs(d, R, K, seed=2)
    b = O.expand_to_batch(inp, K)
    with torch.no_grad():
        ref = model.sample_actions([b["image"]], [torch.ones(R * K, dtype=torch.bool)], b["tokens"], b["masks"],
                                   b["state"], noise=b["noise"].clone())
    masks_r = torch.arange(d.max_lang_len)[None, :] < inp["lens"][:, None]
    ded = O.sample_actions_dedup(w, d, inp["image"], inp["tokens"], masks_r, inp["state"], inp["noise"], K)
    assert (ded - ref).abs().max().item() <= 1e-2   # batch-size dependent bf16 GEMM blocking only


def test_time_embedding_and_schedule_vs_reference():
    """create_sinusoidal_pos_embedding modeling_pi0.py:71-89 and the fp32 time loop :697-714."""
    from oracle import pi0_oracle as O
    M, _ = ref_shim.pi0_modules()
    times, dt = O.denoise_times(10)
    t, ref_times = torch.tensor(1.0, dtype=torch.float32), []
    dtt = torch.tensor(-1.0 / 10, dtype=torch.float32)
    while t >= -dtt / 2:
        ref_times.append(float(t))
        t = t + dtt
    assert times == ref_times and dt == float(dtt)
    for tt in times:
        tv = torch.tensor([tt], dtype=torch.float32)
        ref = M.create_sinusoidal_pos_embedding(tv, 64, min_period=4e-3, max_period=4.0, device=torch.device("cpu"))
        assert torch.equal(O.sinusoidal_time_embedding(tv, 64), ref)


@pytest.mark.parametrize("name,R,K,seed", [("VTINY", 4, 3, 1), ("VTINY", 1, 1, 6), ("VMID", 3, 2, 7),
                                           ("VTINY_MLP", 4, 3, 1), ("VMID_MLP", 3, 2, 7)])
def test_verifier_scores_vs_reference_object(name, R, K, seed):
    """compute_max_similarity_scores_batch of the real EfficientEnsembleMerged (trunk injected, see
    oracle/make_golden_verifier.py) vs the oracle restatement."""
    from oracle import make_golden_verifier as G
    from oracle import verifier_oracle as V
    d = getattr(V, name)
    w = V.make_verifier_weights(d, seed=0)
    inp = V.make_inputs(d, R * K, seed=seed)
    ref = G.run_reference(d, w, inp, K)
    best, idx, scores, means = V.compute_max_similarity_scores(w, d, inp["image"], inp["tokens"], inp["histories"], K)
    assert idx == ref["global_idx"]
    assert abs(best - ref["max_score"]) < 2e-6


@pytest.mark.parametrize("name,R,K", [("VTINY", 4, 3), ("VTINY_MLP", 3, 2)])
def test_reference_slow_path_consumes_pair_zero_only(name, R, K):
    """with DISTINCT (image, instruction) pairs the reference encodes every pair
    (efficient_ensemble_merged.py:348-376), tiles them over the N trajectories (the repeat of :213-214), and still reads
    row 0 of the similarity matrix (:422-425) - the result is the one of pair 0 alone, and the instruction it returns is
    instructions[min(g* . K, len - 1)] (:444).  That is what cvb_verifier_score computes (one context: pair 0) and what the
    EfficientEnsembleMerged mirror returns; shown here on the reference's own code."""
    from oracle import ref_verifier
    from oracle import verifier_oracle as V
    d = getattr(V, name)
    w = V.make_verifier_weights(d, seed=0)
    N = R * K
    inp = V.make_inputs(d, N, seed=11)
    g = torch.Generator().manual_seed(5)
    images = [inp["image"][0]] + [torch.rand(3, d.image, d.image, generator=g) * 2 - 1 for _ in range(N - 1)]
    tokens = [inp["tokens"][0]] + [torch.randint(1, d.vocab - 1, (d.text_ctx,), generator=g) for _ in range(N - 1)]
    calls = []

    def feature_fn(img_tensor, text_tokens):  # the trunk restatement, evaluated on the pair the reference hands over
        calls.append((img_tensor.clone(), text_tokens.clone()))
        return V.extract_features(w, d, img_tensor, text_tokens)

    ens = ref_verifier.build_reference_ensemble(d, w, feature_fn)
    ens.preprocess = lambda im: im
    with torch.no_grad():
        ms, mi, mh, gi = ens.compute_max_similarity_scores_batch(images, tokens, inp["histories"],
                                                                 cfg_repeat_language_instructions=K)
    assert len(calls) == N  # every pair WAS encoded ...
    assert not torch.equal(calls[1][1], calls[0][1])
    # ... and only pair 0 decided
    best, idx, scores, means = V.compute_max_similarity_scores(w, d, inp["image"], inp["tokens"], inp["histories"], K)
    assert int(gi) == idx
    assert abs(float(ms) - best) < 2e-6
    assert mi is tokens[min((idx // K) * K, N - 1)]
    assert mh is inp["histories"][idx]
    # a different pair 0 gives a different decision value: the check above is not vacuous
    with torch.no_grad():
        ms2, *_ = ens.compute_max_similarity_scores_batch(images[1:] + images[:1], tokens[1:] + tokens[:1], inp["histories"],
                                                          cfg_repeat_language_instructions=K)
    assert abs(float(ms2) - best) > 1e-6


def test_state_token_kv_is_step_and_sample_invariant_on_the_reference():
    """SURVEY.md F7 on the reference's own code: the suffix's state token attends the prefix and itself only (suffix
    att mask [1, 1, 0, ...], modeling_pi0.py:590, 619) and its embedding is state_proj(state) alone (:577-580), so its
    K / V rows in every expert layer are identical in all Euler steps and across the K samples of a rephrase - the hoist
    the CUDA path applies (state-token K / V once per rephrase, 4 action rows per candidate in steps 1..9) is exact."""
    from oracle import pi0_oracle as O
    d = O.TINY
    R, K = 2, 3
    model, w = _ref_model(d, 3)
    inp = O.make_inputs(d, R, K, seed=3)
    b = O.expand_to_batch(inp, K)
    expert = model.paligemma_with_expert.gemma_expert.model
    seen = {}
    hooks = []
    for li, layer in enumerate(expert.layers):
        for nm in ("k_proj", "v_proj"):
            def hook(mod, args, out, key=(li, nm)):
                if out.shape[1] == 1 + d.chunk_size:  # the suffix pass (state token + action tokens), not the prefix pass
                    seen.setdefault(key, []).append(out[:, 0].detach().clone())
            hooks.append(getattr(layer.self_attn, nm).register_forward_hook(hook))
    with torch.no_grad():
        model.sample_actions([b["image"]], [torch.ones(R * K, dtype=torch.bool)], b["tokens"], b["masks"], b["state"],
                             noise=b["noise"].clone())
    for h in hooks:
        h.remove()
    assert len(seen) == 2 * d.layers
    for key, rows in seen.items():
        assert len(rows) == d.num_steps, key
        for step_rows in rows[1:]:
            assert torch.equal(step_rows, rows[0]), key          # identical in every Euler step
        r0 = rows[0].reshape(R, K, -1)
        assert torch.equal(r0, r0[:, :1].expand_as(r0)), key     # identical across the K samples of a rephrase
    k_last = seen[(d.layers - 1, "k_proj")][0].reshape(R, K, -1)
    assert not torch.equal(k_last[0], k_last[1])                 # ... and it does depend on the rephrase


@pytest.mark.parametrize("name", ["VTINY", "VMID_MLP"])
def test_gate_call_equals_score_zero_of_the_full_call_on_the_reference(name):
    """SURVEY.md F6 on the reference's own code: the 1-candidate gate call of run_simpler_eval_with_openpi.py:344-352 (same
    image, same instruction, candidate 0 alone) returns the score the N-candidate call of :355-363 gives candidate 0 - one
    pass yields both answers, which is how CoverStep / EpisodeBatchDriver apply the 0.1 gate without a second launch."""
    from oracle import make_golden_verifier as G
    from oracle import verifier_oracle as V
    d = getattr(V, name)
    w = V.make_verifier_weights(d, seed=0)
    R, K = 4, 3
    inp = V.make_inputs(d, R * K, seed=9)
    full = G.run_reference(d, w, inp, K)
    fit = full["it"].mean(0, keepdim=True)
    fit = fit / fit.norm(dim=-1, keepdim=True)
    fact = full["act"].mean(0)
    fact = fact / fact.norm(dim=-1, keepdim=True)
    score0_of_full_call = float((fit @ fact.T)[0, 0])
    gate = G.run_reference(d, w, dict(inp, histories=inp["histories"][:1]), 1)
    assert gate["global_idx"] == 0
    assert abs(gate["max_score"] - score0_of_full_call) < 1e-6


def test_last_prefix_layer_tail_is_dead_code_on_the_reference():
    """DESIGN.md section 1 (de-duplication): the prefix pass keeps only the KV cache (modeling_pi0.py:688-695 discards the
    output embeddings), so the LAST PaliGemma layer's o_proj, post-attention norm, MLP and the final norm never reach the
    sampled actions - the engine stops that layer after K / V.  Shown on the reference: scrambling those weights leaves
    sample_actions bit-identical, scrambling the same layer's k_proj does not."""
    from oracle import pi0_oracle as O
    d = O.TINY
    R, K = 2, 2
    model, w = _ref_model(d, 4)
    inp = O.make_inputs(d, R, K, seed=4)
    b = O.expand_to_batch(inp, K)

    def run():
        with torch.no_grad():
            return model.sample_actions([b["image"]], [torch.ones(R * K, dtype=torch.bool)], b["tokens"], b["masks"],
                                        b["state"], noise=b["noise"].clone())

    ref = run()
    lm = model.paligemma_with_expert.paligemma.language_model.model
    last = lm.layers[d.layers - 1]
    g = torch.Generator().manual_seed(1)
    dead = [last.self_attn.o_proj.weight, last.post_attention_layernorm.weight, last.mlp.gate_proj.weight,
            last.mlp.up_proj.weight, last.mlp.down_proj.weight, lm.norm.weight]
    with torch.no_grad():
        for p in dead:
            p.copy_(torch.randn(p.shape, generator=g).to(p.dtype))
    assert torch.equal(run(), ref)
    with torch.no_grad():
        p = last.self_attn.k_proj.weight
        p.copy_(torch.randn(p.shape, generator=g).to(p.dtype) * 0.05)
    assert not torch.equal(run(), ref)


@pytest.mark.parametrize("name,T", [("VTINY", 10), ("VTINY", 6), ("VMID_MLP", 10)])
def test_predict_and_fuse_embeddings_vs_reference_object(name, T):
    """SURVEY.md section 8 a16 on the reference's own code: EfficientEnsembleMerged.fuse_embeddings / predict
    (efficient_ensemble_merged.py:249-307) for one (image, instruction) and N equal-length histories.  The reference does
    NOT pad here (np.array of the histories, :266-267); the mirror (and cvb_verifier_score) always left-pads to 10 rows
    of -5, which the trajectory transformer masks as keys and excludes from the mean - shown equivalent at T = 6."""
    import numpy as np
    from oracle import ref_verifier
    from oracle import verifier_oracle as V
    d = getattr(V, name)
    w = V.make_verifier_weights(d, seed=0)
    N = 7
    inp = V.make_inputs(d, N, seed=13)
    g = torch.Generator().manual_seed(17)
    hist = []
    for _ in range(N):
        a = torch.rand(T, d.action_dim, generator=g) * 2 - 1
        a[:, :6] *= 0.05
        a[:, 6] = (a[:, 6] > 0).float()
        hist.append(a.numpy())
    patch, text = V.extract_features(w, d, inp["image"], inp["tokens"])
    ens = ref_verifier.build_reference_ensemble(d, w, lambda img, tok: (patch, text))
    ens.preprocess = lambda im: inp["image"][0]
    with torch.no_grad():
        fit, fact = ens.fuse_embeddings(inp["image"][0], inp["tokens"][0], hist)
        best_hist, score_dict = ens.predict(inp["image"][0], inp["tokens"][0], hist)
    assert fit.shape == (N, d.embed) and torch.equal(fit, fit[:1].expand_as(fit))  # N identical image-text rows
    scores = V.scores_from_features(w, d, patch, text, V.pad_histories(hist, d.history))
    ref_scores = torch.tensor([score_dict[str(i)] for i in range(N)])
    assert (scores - ref_scores).abs().max().item() < 2e-6
    assert best_hist is hist[int(scores.argmax())]
    assert np.array_equal(best_hist, hist[int(ref_scores.argmax())])
