"""The verifier trunk restatement (oracle/verifier_oracle.py trunk_image_patches / trunk_text_tokens) against an INDEPENDENT
implementation of the published SigLIP architecture: Hugging Face transformers' SiglipVisionModel / SiglipTextModel
(transformers 5.5.0 in this image; runs on the GPU box too - no /root/reference needed).

Why: the reference's trunk lives in open_clip_torch + timm, which are not vendored and not installed (SURVEY.md section 8c:
"parity unpinned"); the oracle's trunk follows the published timm / open_clip modules from knowledge that cannot be checked
offline.  HF's SigLIP is a separate code base for the same architecture (pre-norm blocks, GELU-tanh MLP, LayerNorm eps 1e-6,
learned absolute position embeddings, no class token, no causal / padding mask in the text tower, a linear head after the
final LayerNorm).  Loading the SAME weights into both and comparing the two hook points the reference reads
(`visual.trunk.blocks[-1].attn` output, finetune_trajectory_bridge_ddp.py:272-274; `text_projection(ln_final(.))` per token,
:320-327) pins the restatement's structure (measured here: bit-identical in fp32 AND in bf16 on this torch build; the gates
leave room for another build's kernels; negative controls - erf GELU, another LayerNorm eps - do differ).  It does not prove that timm /
open_clip agree with HF - the trunk stays "unpinned against the reference's own dependency" - but it removes the risk that the
oracle (and therefore the CUDA trunk checked against it) implements something other than SigLIP."""
import pytest
import torch

from oracle import verifier_oracle as V

transformers = pytest.importorskip("transformers")
TRK = "verifier.trunk."


def _hf_models(d, act="gelu_pytorch_tanh", eps=1e-6):
    from transformers import SiglipTextConfig, SiglipTextModel, SiglipVisionConfig, SiglipVisionModel
    vc = SiglipVisionConfig(hidden_size=d.width, intermediate_size=d.mlp, num_hidden_layers=d.layers,
                            num_attention_heads=d.heads, image_size=d.image, patch_size=d.patch,
                            hidden_act=act, layer_norm_eps=eps)
    tc = SiglipTextConfig(vocab_size=d.vocab, hidden_size=d.width, intermediate_size=d.mlp, num_hidden_layers=d.text_layers,
                          num_attention_heads=d.heads, max_position_embeddings=d.text_ctx, hidden_act=act,
                          layer_norm_eps=eps, projection_size=d.width)
    return SiglipVisionModel(vc).eval(), SiglipTextModel(tc).eval()


def _load(d, w, vis, txt, dtype):
    """The oracle's (timm / open_clip named) trunk weights into the HF modules."""
    W = d.width
    v, t = TRK + "visual.trunk.", TRK + "text."
    sv = {"vision_model.embeddings.patch_embedding.weight": w[v + "patch_embed.proj.weight"],
          "vision_model.embeddings.patch_embedding.bias": w[v + "patch_embed.proj.bias"],
          "vision_model.embeddings.position_embedding.weight": w[v + "pos_embed"][0]}
    for l in range(d.layers):
        p, h = v + f"blocks.{l}.", f"vision_model.encoder.layers.{l}."
        qkv_w, qkv_b = w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"]
        for i, n in enumerate(("q_proj", "k_proj", "v_proj")):
            sv[h + f"self_attn.{n}.weight"], sv[h + f"self_attn.{n}.bias"] = qkv_w[i * W:(i + 1) * W], qkv_b[i * W:(i + 1) * W]
        sv[h + "self_attn.out_proj.weight"], sv[h + "self_attn.out_proj.bias"] = w[p + "attn.proj.weight"], w[p + "attn.proj.bias"]
        sv[h + "layer_norm1.weight"], sv[h + "layer_norm1.bias"] = w[p + "norm1.weight"], w[p + "norm1.bias"]
        if l < d.layers - 1:  # the last block's MLP is never read (the hook sits on its attention)
            sv[h + "layer_norm2.weight"], sv[h + "layer_norm2.bias"] = w[p + "norm2.weight"], w[p + "norm2.bias"]
            for a, b in (("fc1", "fc1"), ("fc2", "fc2")):
                sv[h + f"mlp.{b}.weight"], sv[h + f"mlp.{b}.bias"] = w[p + f"mlp.{a}.weight"], w[p + f"mlp.{a}.bias"]
    st = {"text_model.embeddings.token_embedding.weight": w[t + "token_embedding.weight"],
          "text_model.embeddings.position_embedding.weight": w[t + "positional_embedding"],
          "text_model.final_layer_norm.weight": w[t + "ln_final.weight"], "text_model.final_layer_norm.bias": w[t + "ln_final.bias"],
          "text_model.head.weight": w[t + "text_projection.weight"], "text_model.head.bias": w[t + "text_projection.bias"]}
    for l in range(d.text_layers):
        p, h = t + f"transformer.resblocks.{l}.", f"text_model.encoder.layers.{l}."
        qkv_w, qkv_b = w[p + "attn.in_proj_weight"], w[p + "attn.in_proj_bias"]
        for i, n in enumerate(("q_proj", "k_proj", "v_proj")):
            st[h + f"self_attn.{n}.weight"], st[h + f"self_attn.{n}.bias"] = qkv_w[i * W:(i + 1) * W], qkv_b[i * W:(i + 1) * W]
        st[h + "self_attn.out_proj.weight"], st[h + "self_attn.out_proj.bias"] = w[p + "attn.out_proj.weight"], w[p + "attn.out_proj.bias"]
        st[h + "layer_norm1.weight"], st[h + "layer_norm1.bias"] = w[p + "ln_1.weight"], w[p + "ln_1.bias"]
        st[h + "layer_norm2.weight"], st[h + "layer_norm2.bias"] = w[p + "ln_2.weight"], w[p + "ln_2.bias"]
        st[h + "mlp.fc1.weight"], st[h + "mlp.fc1.bias"] = w[p + "mlp.c_fc.weight"], w[p + "mlp.c_fc.bias"]
        st[h + "mlp.fc2.weight"], st[h + "mlp.fc2.bias"] = w[p + "mlp.c_proj.weight"], w[p + "mlp.c_proj.bias"]
    missing_v = vis.load_state_dict({k: x.to(dtype) for k, x in sv.items()}, strict=False)
    missing_t = txt.load_state_dict({k: x.to(dtype) for k, x in st.items()}, strict=True)
    # the only HF vision tensors without a counterpart: the unread last-block MLP / norm2, post_layernorm and the pooling head
    last = f"vision_model.encoder.layers.{d.layers - 1}."
    assert all(k.startswith((last + "mlp.", last + "layer_norm2.", "vision_model.post_layernorm.", "vision_model.head."))
               for k in missing_v.missing_keys), missing_v.missing_keys
    assert not missing_v.unexpected_keys and not missing_t.missing_keys and not missing_t.unexpected_keys
    return vis.to(dtype), txt.to(dtype)


def _hf_features(vis, txt, image, tokens):
    grabbed = {}
    hook = vis.vision_model.encoder.layers[-1].self_attn.register_forward_hook(
        lambda mod, args, out: grabbed.__setitem__("attn", out[0] if isinstance(out, tuple) else out))
    with torch.no_grad():
        vis(pixel_values=image.to(next(vis.parameters()).dtype))
        hidden = txt(input_ids=tokens).last_hidden_state          # = final_layer_norm(encoder output), every token
        text = txt.text_model.head(hidden)
    hook.remove()
    return grabbed["attn"], text


@pytest.mark.parametrize("name", ["VTINY", "VMID"])
def test_trunk_restatement_equals_hf_siglip_in_fp32(name):
    d = getattr(V, name)
    w = V.make_verifier_weights(d, seed=4)
    inp = V.make_inputs(d, 1, seed=4)
    vis, txt = _load(d, w, *_hf_models(d), torch.float32)
    hf_patch, hf_text = _hf_features(vis, txt, inp["image"], inp["tokens"])
    with V.truth_mode():
        tw = V.truth_weights(w)
        patch = V.trunk_image_patches(tw, d, inp["image"])
        text = V.trunk_text_tokens(tw, d, inp["tokens"])
    assert patch.shape == hf_patch.shape == (1, d.n_patches, d.width) and text.shape == hf_text.shape == (1, d.text_ctx, d.width)
    for ours, hf, what in ((patch, hf_patch, "patch hook"), (text, hf_text, "text tokens")):
        rel = ((ours - hf).norm() / hf.norm()).item()
        print(f"{name} {what}: rel-L2 vs HF SigLIP (fp32) {rel:.2e}")
        assert rel < 2e-5, (what, rel)
        assert float(hf.abs().mean()) > 1e-3  # not a comparison of zeros
    # negative controls: the comparison does see the architecture choices the restatement had to guess
    for kw in (dict(act="gelu"), dict(eps=1e-5)):
        vis2, txt2 = _load(d, w, *_hf_models(d, **kw), torch.float32)
        p2, t2 = _hf_features(vis2, txt2, inp["image"], inp["tokens"])
        assert ((patch - p2).norm() / p2.norm()).item() > 1e-6 and ((text - t2).norm() / t2.norm()).item() > 1e-6, kw


def test_trunk_restatement_matches_hf_siglip_in_bf16():
    """The dtype the reference runs the trunk in (efficient_ensemble_merged.py:66): both sides round after every op, the op
    order may differ in places (fused qkv vs three projections), so agreement is to bf16 noise, not to the bit."""
    d = V.VMID
    w = V.make_verifier_weights(d, seed=4)
    inp = V.make_inputs(d, 1, seed=4)
    vis, txt = _load(d, w, *_hf_models(d), torch.bfloat16)
    hf_patch, hf_text = _hf_features(vis, txt, inp["image"], inp["tokens"])
    patch = V.trunk_image_patches(w, d, inp["image"])
    text = V.trunk_text_tokens(w, d, inp["tokens"])
    for ours, hf, what in ((patch, hf_patch, "patch hook"), (text, hf_text, "text tokens")):
        rel = ((ours.float() - hf.float()).norm() / hf.float().norm()).item()
        print(f"VMID {what}: rel-L2 vs HF SigLIP (bf16) {rel:.2e}")
        assert rel < 2e-2, (what, rel)
