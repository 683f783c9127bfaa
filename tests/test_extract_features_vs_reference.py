"""The reference's own `VLA_SigLIP2_Bridge.extract_features` (bridge_verifier/ensemble_eval/finetune_trajectory_bridge_ddp.py
:297-355) EXECUTED UNMODIFIED on an open_clip-shaped model whose towers are Hugging Face transformers' SigLIP (an independent
implementation of the published architecture; open_clip / timm are absent offline) carrying the oracle's trunk weights.

What this pins (CPU, authoring container): everything the reference does AROUND the third-party trunk - the bf16 cast of the
image, the two hook points (`visual.trunk.blocks[-1].attn` output, `text.transformer` output, :271-278), `ln_final` +
`text_projection` applied to every token in bf16, the class-token rule (:340-349), the fp32 cast and the per-token L2
normalisation - against `oracle.verifier_oracle.extract_features`, which the CUDA trunk is gated against.  Together with
tests/test_trunk_vs_hf_siglip.py (same towers vs the restated trunk) the only unpinned piece left is whether timm / open_clip
deviate from the published SigLIP."""
from types import SimpleNamespace

import pytest
import torch

from oracle import ref_shim
from oracle import verifier_oracle as V
from tests.test_trunk_vs_hf_siglip import _hf_models, _load

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present")


class _Tap(torch.nn.Module):
    """Runs `inner` and passes the tensor the open_clip module would return through `tap` (an Identity): a forward hook on
    `tap` sees exactly what the reference's hook sees on the open_clip / timm module (a tensor, not HF's tuple / dataclass)."""

    def __init__(self, inner, pick):
        super().__init__()
        self.inner, self.tap, self.pick = inner, torch.nn.Identity(), pick

    def forward(self, *a, **kw):
        out = self.inner(*a, **kw)
        self.tap(self.pick(out))
        return out


def _open_clip_shaped(vis, txt, with_class_token=False):
    last = vis.vision_model.encoder.layers[-1]
    last.self_attn = _Tap(last.self_attn, lambda out: out[0] if isinstance(out, tuple) else out)
    txt.text_model.encoder = _Tap(txt.text_model.encoder, lambda out: out.last_hidden_state)
    attn_tap = last.self_attn.tap
    if with_class_token:  # a trunk WITH a class token (577 tokens): the reference drops the first one, :340-342
        class Prepend(torch.nn.Module):
            def forward(self, x):
                return torch.cat([torch.full_like(x[:, :1], 7.0), x], dim=1)
        attn_tap = Prepend()
        last.self_attn.tap = attn_tap
    model = SimpleNamespace(
        visual=SimpleNamespace(trunk=SimpleNamespace(blocks=[SimpleNamespace(attn=attn_tap)])),
        text=SimpleNamespace(transformer=txt.text_model.encoder.tap, ln_final=txt.text_model.final_layer_norm,
                             text_projection=txt.text_model.head),
        encode_text=lambda text, normalize=False: txt(input_ids=text).pooler_output,
        encode_image=lambda images, normalize=False: vis(pixel_values=images).pooler_output)
    return model


@pytest.mark.parametrize("name,cls_token", [("VTINY", False), ("VMID", False), ("VTINY", True)])
def test_reference_extract_features_unmodified_vs_oracle(name, cls_token):
    _, EM = ref_shim.verifier_modules()
    d = getattr(V, name)
    w = V.make_verifier_weights(d, seed=4)
    inp = V.make_inputs(d, 1, seed=4)
    vis, txt = _load(d, w, *_hf_models(d), torch.bfloat16)  # the reference keeps the frozen encoder in bf16 (:66)
    model = _open_clip_shaped(vis, txt, cls_token)
    # the hook registration of VLA_SigLIP2_Bridge.__init__ (:265-278), on the same attribute paths
    me = SimpleNamespace(model=model, activation={}, num_img_patches=d.n_patches)

    def get_activation(nm):
        def hook(module, inputs, output):
            me.activation[nm] = output
        return hook

    hooks = [model.visual.trunk.blocks[-1].attn.register_forward_hook(get_activation("image_patches")),
             model.text.transformer.register_forward_hook(get_activation("text_features"))]
    with torch.no_grad():
        patch_ref, text_ref = EM.VLA_SigLIP2_Bridge.extract_features(me, inp["image"], inp["tokens"])  # the reference's lines
    for h in hooks:
        h.remove()
    patch, text = V.extract_features(w, d, inp["image"], inp["tokens"])
    assert patch_ref.dtype == text_ref.dtype == torch.float32
    assert patch_ref.shape == patch.shape == (1, d.n_patches, d.width) and text_ref.shape == text.shape == (1, d.text_ctx, d.width)
    for ours, ref, what in ((patch, patch_ref, "patch features"), (text, text_ref, "text features")):
        rel = ((ours - ref).norm() / ref.norm()).item()
        print(f"{name} {what}: rel-L2 oracle vs the reference's extract_features over HF SigLIP (bf16 trunk) {rel:.2e}")
        assert rel < 2e-2, (what, rel)  # bf16 trunk: measured 0 on this torch build, the gate leaves room for another build
        assert torch.allclose(ref.norm(dim=-1), torch.ones(ref.shape[:2]), atol=1e-5)  # unit rows, as the heads expect


def test_whole_reference_decision_without_any_oracle_piece():
    """End to end with NO oracle code on the reference side: HF SigLIP towers -> the reference's extract_features -> the
    reference's EfficientEnsembleMerged.compute_max_similarity_scores_batch (heads, fusion, selection) - against the
    oracle's compute_max_similarity_scores on the same weights and inputs.  (The committed verifier goldens were made with the
    oracle's trunk injected into the reference object; this shows that choice changes nothing.)"""
    from oracle import ref_verifier
    _, EM = ref_shim.verifier_modules()
    d = V.VMID
    R, K = 3, 2
    w = V.make_verifier_weights(d, seed=0)
    inp = V.make_inputs(d, R * K, seed=21)
    vis, txt = _load(d, w, *_hf_models(d), torch.bfloat16)
    model = _open_clip_shaped(vis, txt)
    me = SimpleNamespace(model=model, activation={}, num_img_patches=d.n_patches)
    model.visual.trunk.blocks[-1].attn.register_forward_hook(lambda m, i, o: me.activation.__setitem__("image_patches", o))
    model.text.transformer.register_forward_hook(lambda m, i, o: me.activation.__setitem__("text_features", o))
    ens = ref_verifier.build_reference_ensemble(
        d, w, lambda img, tok: EM.VLA_SigLIP2_Bridge.extract_features(me, img, tok))
    ens.preprocess = lambda im: inp["image"][0]
    N = R * K
    with torch.no_grad():
        ms, mi, mh, gi = ens.compute_max_similarity_scores_batch([inp["image"][0]] * N, [inp["tokens"][0]] * N,
                                                                 inp["histories"], cfg_repeat_language_instructions=K)
    best, idx, scores, means = V.compute_max_similarity_scores(w, d, inp["image"], inp["tokens"], inp["histories"], K)
    assert int(gi) == idx
    assert abs(float(ms) - best) < 2e-6


class _OpenClipShapedSigLIP(torch.nn.Module):
    """What open_clip.create_model_from_pretrained('hf-hub:timm/ViT-L-16-SigLIP2-384') returns, as far as the reference
    touches it (efficient_ensemble_merged.py:57-89, finetune_trajectory_bridge_ddp.py:183-214, 265-278, 297-355): an
    nn.Module with .visual.trunk.{num_features, patch_embed.proj.kernel_size, blocks[-1].attn}, .visual.image_size,
    .text.{output_dim, transformer, ln_final, text_projection}, encode_image / encode_text - over HF SigLIP towers."""

    def __init__(self, vis, txt, d):
        super().__init__()
        self.vis_hf, self.txt_hf = vis, txt
        self.context_length = d.text_ctx  # read by the string-instruction branches (:256-257, :342)
        last = vis.vision_model.encoder.layers[-1]
        last.self_attn = _Tap(last.self_attn, lambda out: out[0] if isinstance(out, tuple) else out)
        txt.text_model.encoder = _Tap(txt.text_model.encoder, lambda out: out.last_hidden_state)
        self.visual = SimpleNamespace(
            image_size=(d.image, d.image),
            trunk=SimpleNamespace(num_features=d.width, blocks=[SimpleNamespace(attn=last.self_attn.tap)],
                                  patch_embed=SimpleNamespace(proj=SimpleNamespace(kernel_size=(d.patch, d.patch)))))
        self.text = SimpleNamespace(output_dim=d.width, transformer=txt.text_model.encoder.tap,
                                    ln_final=txt.text_model.final_layer_norm, text_projection=txt.text_model.head)

    def encode_text(self, text, normalize=False):
        return self.txt_hf(input_ids=text).pooler_output

    def encode_image(self, images, normalize=False):
        return self.vis_hf(pixel_values=images).pooler_output


def _merged_checkpoint(v, vw):
    """The merged .pt layout of the reference (one dict of state dicts per member, :94-184) from the synthetic weights -
    the same construction tests/test_loaders_gpu.py feeds the EfficientEnsembleMerged MIRROR."""
    comps = []
    for m in range(v.members):
        c = {}
        pre = f"verifier.{m}."
        for k, t in vw.items():
            if k.startswith(pre):
                comp, nm = k[len(pre):].split(".", 1)
                c.setdefault(comp, {})[nm] = t
        c["action_padding_value"] = -5.0
        for unused in ("single_step_action_encoder", "trajectory_encoder", "complex_action_encoder"):
            c.setdefault(unused, None)
        comps.append(c)
    return {"ensemble_components": comps, "backbone": "hf-hub:timm/ViT-L-16-SigLIP2-384", "use_transformer": v.traj_layers > 0,
            "history_length": v.history, "action_dim": v.action_dim, "num_models": v.members}


@pytest.mark.parametrize("name,weights_only", [("VMID", False), ("VMID", True), ("VMID_MLP", False)])
def test_reference_constructor_and_decision_unmodified(tmp_path, monkeypatch, name, weights_only):
    """The reference's EfficientEnsembleMerged(merged_checkpoint_path, device) - its REAL __init__ (:26-186: torch.load of
    the merged file, strict load_state_dict of every component, VLA_SigLIP2_Bridge with its hooks) and its REAL
    compute_max_similarity_scores_batch / extract_shared_features - on a merged checkpoint in the layout the mirror's loader
    test uses; only open_clip.create_model_from_pretrained / get_tokenizer (absent offline) are replaced, by HF SigLIP
    towers with the oracle's trunk weights.  VMID has the head sizes the reference hard-codes (512 / 8 heads / 4 layers /
    feed-forward 1024 / hidden 512), so nothing else is touched.  Result against the oracle's decision."""
    _, EM = ref_shim.verifier_modules()
    d = getattr(V, name)
    w = V.make_verifier_weights(d, seed=0)
    ck = _merged_checkpoint(d, w)
    if weights_only:  # :40-48 - a checkpoint without metadata takes the CoVer-BridgeV2 defaults
        ck = {"ensemble_components": ck["ensemble_components"]}
    path = tmp_path / "merged.pt"
    torch.save(ck, path)
    created = []

    def create_model_from_pretrained(backbone):
        created.append(backbone)
        vis, txt = _load(d, w, *_hf_models(d), torch.float32)  # the reference casts the encoder to bf16 itself (:66)
        return _OpenClipShapedSigLIP(vis, txt, d), (lambda im: im)

    monkeypatch.setattr(EM, "create_model_from_pretrained", create_model_from_pretrained)
    monkeypatch.setattr(EM, "get_tokenizer", lambda backbone: (lambda texts, context_length=64: None))
    ens = EM.EfficientEnsembleMerged(str(path), device="cpu")
    assert created == ["hf-hub:timm/ViT-L-16-SigLIP2-384"]
    assert (ens.num_models, ens.history_length, ens.action_dim, ens.use_transformer) == (d.members, 10, 7, d.traj_layers > 0)
    assert all(p.dtype == torch.bfloat16 for p in ens.siglip_model.parameters())
    R, K = 3, 2
    N = R * K
    inp = V.make_inputs(d, N, seed=23)
    ms, mi, mh, gi = ens.compute_max_similarity_scores_batch([inp["image"][0]] * N, [inp["tokens"][0]] * N, inp["histories"],
                                                             cfg_repeat_language_instructions=K)
    best, idx, scores, means = V.compute_max_similarity_scores(w, d, inp["image"], inp["tokens"], inp["histories"], K)
    assert int(gi) == idx
    assert abs(float(ms) - best) < 2e-6
    patch_ref, text_ref = ens.extract_shared_features(inp["image"], inp["tokens"])
    patch, text = V.extract_features(w, d, inp["image"], inp["tokens"])
    assert ((patch - patch_ref).norm() / patch_ref.norm()).item() < 2e-2 and ((text - text_ref).norm() / text_ref.norm()).item() < 2e-2


@pytest.mark.skipif(bool(__import__("os").environ.get("CVB_FAST_TESTS")), reason="full-size ViT-L on the CPU (about a minute); CVB_FAST_TESTS=1 skips it")
def test_full_size_golden_is_reproduced_by_the_unmodified_reference_constructor(tmp_path, monkeypatch):
    """BASELINE.json configs[0] at FULL size: tests/golden/verifier_vfull_R8K5.pt (made with the oracle's trunk injected into the
    reference object) against the reference's real constructor + string-instruction fast path (:330-347) over HF SigLIP
    ViT-L/16-384 towers carrying the same weights - the fixture the CUDA path is gated against at the configuration of record
    does not depend on the oracle's trunk restatement."""
    from pathlib import Path
    _, EM = ref_shim.verifier_modules()
    fx = torch.load(Path(__file__).resolve().parent / "golden" / "verifier_vfull_R8K5.pt")
    d = V.VFULL
    R, K = fx["R"], fx["K"]
    N = R * K
    w = V.make_verifier_weights(d, seed=0)
    inp = V.make_inputs(d, N, seed=fx["seed"])
    path = tmp_path / "merged.pt"
    torch.save(_merged_checkpoint(d, w), path)

    def create_model_from_pretrained(backbone):
        vis, txt = _load(d, w, *_hf_models(d), torch.float32)
        return _OpenClipShapedSigLIP(vis, txt, d), (lambda im: im)

    monkeypatch.setattr(EM, "create_model_from_pretrained", create_model_from_pretrained)
    monkeypatch.setattr(EM, "get_tokenizer", lambda backbone: (lambda texts, context_length=64: inp["tokens"][:1].clone()))
    ens = EM.EfficientEnsembleMerged(str(path), device="cpu")
    instr = "put the carrot on the plate"
    ms, mi, mh, gi = ens.compute_max_similarity_scores_batch([inp["image"][0]] * N, [instr] * N, inp["histories"],
                                                             cfg_repeat_language_instructions=K)
    print(f"VFULL R{R}K{K}: reference constructor path max_score {float(ms):.9f} idx {int(gi)}; golden {fx['max_score']:.9f} idx {fx['global_idx']}")
    assert mi == instr and int(gi) == fx["global_idx"]
    assert abs(float(ms) - fx["max_score"]) <= 1e-3 * float(fx["scores"].abs().max())
