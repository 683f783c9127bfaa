"""The two constructors run_simpler_eval_with_openpi.py actually calls - PI0Policy.from_pretrained(path) (:149) and
EfficientEnsembleMerged(ckpt_path, device) (:164) - on synthetic checkpoints written in the reference's formats, plus the
secondary verifier surfaces (fuse_embeddings, extract_shared_features) against the oracle."""
import json
import sys
import types

import numpy as np
import pytest
import torch

from oracle import pi0_oracle as O
from oracle import verifier_oracle as V
from tests.helpers import SCORE_TOL, action_gate, max_abs, pi0_truth, rel_l2, score_gate, verifier_truth_scores

pytestmark = pytest.mark.gpu


def _write_pi0_checkpoint(path, d, w, norm="IDENTITY", with_stats=False):
    """config.json + model.safetensors as PreTrainedPolicy.save_pretrained writes them (policies/pretrained.py:71-151)."""
    from safetensors.torch import save_file
    cfg = dict(type="pi0", chunk_size=d.chunk_size, n_action_steps=d.chunk_size, tokenizer_max_length=d.max_lang_len,
               proj_width=d.ex_width, num_steps=d.num_steps, max_state_dim=d.max_state_dim, max_action_dim=d.max_action_dim,
               resize_imgs_with_padding=[d.vis_image, d.vis_image],
               input_features={"observation.images.top": {"type": "VISUAL", "shape": [3, d.vis_image, d.vis_image]},
                               "observation.state": {"type": "STATE", "shape": [7]}},
               output_features={"action": {"type": "ACTION", "shape": [7]}},
               normalization_mapping={"VISUAL": "IDENTITY", "STATE": norm, "ACTION": "IDENTITY"},
               # model dimensions are configuration here (the reference hard-codes the 3.5 B model)
               vis_layers=d.vis_layers, vis_width=d.vis_width, vis_heads=d.vis_heads, vis_mlp=d.vis_mlp, vis_patch=d.vis_patch,
               vis_image=d.vis_image, layers=d.layers, lm_width=d.lm_width, lm_mlp=d.lm_mlp, heads=d.heads,
               head_dim=d.head_dim, ex_mlp=d.ex_mlp, vocab=d.vocab, max_rephrases=3, max_samples=2,
               an_unknown_key_future_versions_may_add=1)
    (path / "config.json").write_text(json.dumps(cfg))
    # published pi0 safetensors are stored in float32: the loader must cast like the reference's load_state_dict does
    sd = {"model." + k: t.to(torch.float32).contiguous() for k, t in w.items()}
    sd["model.paligemma_with_expert.paligemma.language_model.lm_head.weight"] = torch.zeros(8, d.lm_width)  # unused by sampling
    sd["model.paligemma_with_expert.gemma_expert.lm_head.weight"] = torch.zeros(8, d.ex_width)
    if with_stats:
        sd["normalize_inputs.buffer_observation_state.mean"] = torch.full((7,), 0.25)
        sd["normalize_inputs.buffer_observation_state.std"] = torch.full((7,), 2.0)
    save_file(sd, str(path / "model.safetensors"))


def test_pi0_from_pretrained_directory(tmp_path):
    from cover_vla_b200.pi0 import PI0Policy
    d, R, K = O.TINY, 3, 2
    N = R * K
    w = O.make_pi0_weights(d, seed=8)
    _write_pi0_checkpoint(tmp_path, d, w)
    policy = PI0Policy.from_pretrained(tmp_path)
    assert policy.config.chunk_size == d.chunk_size and policy.config.input_features["observation.state"].shape == (7,)
    # lm_head and friends never reach the GPU
    assert not any("lm_head" in k for k in policy.engine._keep)
    inp = O.make_inputs(d, R, K, seed=8)
    b = O.expand_to_batch(inp, K)
    obs = {"observation.images.top": b["image"].cuda(), "observation.state": b["state"][:, :7].cuda(),
           "lang_tokens": b["tokens"].cuda(), "lang_masks": b["masks"].cuda(), "task": ["x"] * N}
    q = policy.select_action(obs, noise=b["noise"].cuda())
    got = torch.stack(list(q), dim=1).cpu()
    ref = O.sample_actions(w, d, b["image"], b["tokens"], b["masks"], b["state"], b["noise"])
    action_gate(got, ref[:, :, :7], pi0_truth(O, w, d, inp, K)[:, :, :7], "from_pretrained TINY")
    policy.engine.close()


def test_pi0_from_pretrained_normalisation(tmp_path):
    """A checkpoint whose config normalises a feature must carry (or be given) the statistics - never a silent IDENTITY."""
    from cover_vla_b200.pi0 import PI0Policy
    d = O.TINY
    w = O.make_pi0_weights(d, seed=8)
    _write_pi0_checkpoint(tmp_path, d, w, norm="MEAN_STD", with_stats=False)
    with pytest.raises(ValueError, match="statistics"):
        PI0Policy.from_pretrained(tmp_path)
    _write_pi0_checkpoint(tmp_path, d, w, norm="MEAN_STD", with_stats=True)
    policy = PI0Policy.from_pretrained(tmp_path)
    R, K = 2, 2
    inp = O.make_inputs(d, R, K, seed=9)
    b = O.expand_to_batch(inp, K)
    raw_state = b["state"][:, :7] * 2.0 + 0.25  # the policy normalises it back to b["state"]
    obs = {"observation.images.top": b["image"].cuda(), "observation.state": raw_state.cuda(),
           "lang_tokens": b["tokens"].cuda(), "lang_masks": b["masks"].cuda(), "task": ["x"] * (R * K)}
    got = torch.stack(list(policy.select_action(obs, noise=b["noise"].cuda())), dim=1).cpu()
    st = b["state"].clone()
    st[:, :7] = (raw_state - 0.25) / (2.0 + 1e-8)
    ref = O.sample_actions(w, d, b["image"], b["tokens"], b["masks"], st, b["noise"])
    inp2 = dict(inp, state=st[:1])
    action_gate(got, ref[:, :, :7], pi0_truth(O, w, d, inp2, K)[:, :, :7], "from_pretrained MEAN_STD")
    policy.engine.close()


def test_adjacent_identical_rephrases_do_not_raise():
    """Two adjacent rephrases that tokenise identically break the R x K grouping; the reference handles any batch."""
    from cover_vla_b200.pi0 import PI0Config, PI0Policy, PolicyFeature
    d, R, K = O.TINY, 3, 2
    w = O.make_pi0_weights(d, 0)
    cfg = PI0Config(chunk_size=d.chunk_size, n_action_steps=d.chunk_size, tokenizer_max_length=d.max_lang_len,
                    proj_width=d.ex_width, num_steps=d.num_steps, vis_layers=d.vis_layers, vis_width=d.vis_width,
                    vis_heads=d.vis_heads, vis_mlp=d.vis_mlp, vis_patch=d.vis_patch, vis_image=d.vis_image,
                    layers=d.layers, lm_width=d.lm_width, lm_mlp=d.lm_mlp, heads=d.heads, head_dim=d.head_dim,
                    ex_mlp=d.ex_mlp, vocab=d.vocab, max_rephrases=R, max_samples=K,
                    resize_imgs_with_padding=(d.vis_image, d.vis_image),
                    input_features={"observation.images.top": PolicyFeature("VISUAL", (3, d.vis_image, d.vis_image)),
                                    "observation.state": PolicyFeature("STATE", (7,))})
    policy = PI0Policy(cfg, state_dict={"model." + k: t for k, t in w.items()})
    inp = O.make_inputs(d, R, K, seed=12)
    inp["tokens"][1], inp["masks"][1], inp["lens"][1] = inp["tokens"][0], inp["masks"][0], inp["lens"][0]  # duplicate
    b = O.expand_to_batch(inp, K)
    obs = {"observation.images.top": b["image"].cuda(), "observation.state": b["state"][:, :7].cuda(),
           "lang_tokens": b["tokens"].cuda(), "lang_masks": b["masks"].cuda(), "task": ["x"] * (R * K)}
    got = torch.stack(list(policy.select_action(obs, noise=b["noise"].cuda())), dim=1).cpu()
    ref = O.sample_actions(w, d, b["image"], b["tokens"], b["masks"], b["state"], b["noise"])
    action_gate(got, ref[:, :, :7], pi0_truth(O, w, d, inp, K)[:, :, :7], "duplicate rephrases")
    policy.engine.close()


def _fake_open_clip(v, vw, calls):
    """Stands in for open_clip_torch (absent here): same three entry points the reference uses
    (efficient_ensemble_merged.py:57-69), serving the synthetic trunk."""
    mod = types.ModuleType("open_clip")

    class _Model:
        def state_dict(self):
            return {k[len("verifier.trunk."):]: t for k, t in vw.items() if k.startswith("verifier.trunk.")}

    def create_model_from_pretrained(name):
        calls.append(("create", name))
        from cover_vla_b200.verifier.efficient_ensemble_merged import default_preprocess
        return _Model(), default_preprocess(v.image)

    def get_tokenizer(name):
        calls.append(("tokenizer", name))

        def tok(texts, context_length=64):
            out = torch.zeros(len(texts), context_length, dtype=torch.int64)
            for i, t in enumerate(texts):
                ids = [1 + (ord(c) * 7 + j) % (v.vocab - 2) for j, c in enumerate(t)][:context_length]
                out[i, : len(ids)] = torch.tensor(ids)
            return out
        return tok

    mod.create_model_from_pretrained = create_model_from_pretrained
    mod.get_tokenizer = get_tokenizer
    return mod


@pytest.mark.parametrize("vname", ["VTINY", "VMID_MLP"])
def test_ensemble_from_merged_checkpoint_path(tmp_path, monkeypatch, vname):
    """merged .pt -> EfficientEnsembleMerged(path): transformer (use_transformer = True, :135-160) and MLP action encoder
    (use_transformer = False, :161-171) checkpoints."""
    from cover_vla_b200.verifier import EfficientEnsembleMerged
    v = getattr(V, vname)
    vw = V.make_verifier_weights(v, 0)
    comps = []
    for m in range(v.members):
        c = {}
        for k, t in vw.items():
            pre = f"verifier.{m}."
            if k.startswith(pre):
                comp, name = k[len(pre):].split(".", 1)
                c.setdefault(comp, {})[name] = t
        c["action_padding_value"] = -5.0
        for unused in ("single_step_action_encoder", "trajectory_encoder", "complex_action_encoder"):
            c.setdefault(unused, None)  # the reference stores None for the encoder the checkpoint does not use
        comps.append(c)
    ckpt = tmp_path / "ensemble.pt"
    torch.save({"ensemble_components": comps, "backbone": "hf-hub:timm/ViT-L-16-SigLIP2-384", "use_transformer": v.traj_layers > 0,
                "history_length": v.history, "action_dim": v.action_dim, "num_models": v.members}, ckpt)
    vf_cfg = dict(vf_image=v.image, vf_patch=v.patch, vf_width=v.width, vf_layers=v.layers, vf_heads=v.heads,
                  vf_mlp=v.mlp, vf_text_layers=v.text_layers, vf_text_ctx=v.text_ctx, vf_vocab=v.vocab,
                  vf_embed=v.embed, vf_pool_heads=v.pool_heads, vf_pool_layers=v.pool_layers)
    if v.traj_layers > 0:  # MLP checkpoints: the mirror reads the encoder type and hidden width from the checkpoint
        vf_cfg.update(vf_traj_layers=v.traj_layers, vf_traj_ff=v.traj_ff)
    # without open_clip the constructor says exactly what is missing
    monkeypatch.setitem(sys.modules, "open_clip", None)
    with pytest.raises(RuntimeError, match="open_clip"):
        EfficientEnsembleMerged(str(ckpt), device="cuda", vf_config=vf_cfg)
    calls = []
    monkeypatch.setitem(sys.modules, "open_clip", _fake_open_clip(v, vw, calls))
    R, K = 3, 2
    N = R * K
    ens = EfficientEnsembleMerged(str(ckpt), device="cuda", vf_config=vf_cfg, max_candidates=N)
    assert calls == [("create", "hf-hub:timm/ViT-L-16-SigLIP2-384"), ("tokenizer", "hf-hub:timm/ViT-L-16-SigLIP2-384")]
    assert ens.num_models == v.members and ens.history_length == v.history and ens.action_dim == v.action_dim
    # string instructions through the tokenizer, PIL frames through the transform - the caller's exact usage (:357-363)
    from PIL import Image
    frame = np.random.default_rng(3).integers(0, 256, size=(v.image, v.image, 3)).astype(np.uint8)
    vin = V.make_inputs(v, N, seed=4)
    instr = "put the spoon on the towel"
    ms, mi, mh, gi = ens.compute_max_similarity_scores_batch([Image.fromarray(frame)] * N, [instr] * N, vin["histories"],
                                                             cfg_repeat_language_instructions=K)
    tok = ens.tokenizer([instr], context_length=v.text_ctx)
    img = ens.preprocess(Image.fromarray(frame))[None]
    best, idx, ref_scores, means = V.compute_max_similarity_scores(vw, v, img, tok, vin["histories"], K)
    traj = V.pad_histories(vin["histories"], v.history)
    err = score_gate(ens.last_scores.cpu(), ref_scores, verifier_truth_scores(V, vw, v, img, tok, traj), "merged checkpoint")
    assert mi == instr and abs(ms - best) <= max(SCORE_TOL, 2 * err)
    # secondary surfaces (:188-192, :249-293)
    patch, text = ens.extract_shared_features(img, tok)
    pr, tr = V.extract_features(vw, v, img, tok)
    assert patch.shape == (1, v.n_patches, v.width) and text.shape == (1, v.text_ctx, v.width)
    assert rel_l2(patch, pr) < 3e-2 and rel_l2(text, tr) < 3e-2
    fit, fact = ens.fuse_embeddings(Image.fromarray(frame), instr, vin["histories"])
    assert fit.shape == (N, v.embed) and fact.shape == (N, v.embed)
    its = torch.stack([V.image_text_embedding(vw, m, v, pr, tr)[0] for m in range(v.members)]).mean(0)
    its = its / its.norm()
    acts = torch.stack([V.trajectory_embedding(vw, m, v, traj) for m in range(v.members)]).mean(0)
    acts = acts / acts.norm(dim=-1, keepdim=True)
    assert rel_l2(fit[0], its) < 5e-3 and torch.equal(fit[0], fit[N - 1])
    assert rel_l2(fact, acts) < 1e-4  # the trajectory side is fp32 on both sides
    assert max_abs((fit * fact).sum(-1), ens.last_scores) < 1e-5
    ens.engine.close()


def test_two_devices_one_process():
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per device: a second handle on another GPU of the same process must
    get it too (round-1 advisor finding)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests.helpers import build_pi0_engine
    d, R, K = O.MID, 2, 2
    w = O.make_pi0_weights(d, 0)
    inp = O.make_inputs(d, R, K, seed=1)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        eng = build_pi0_engine(d, w, R, K, device=dev)
        with torch.cuda.device(dev):
            args = (inp["image"][0].to(dev).contiguous(), inp["tokens"].to(dev), inp["lens"].to(torch.int32).to(dev),
                    inp["state"][0].to(dev).contiguous(), inp["noise"].to(dev))
            outs.append(eng.pi0_sample(*args, K=K).cpu())
            torch.cuda.synchronize(dev)
        eng.close()
    assert torch.equal(outs[0], outs[1])
