"""TEST INFRASTRUCTURE - generates tests/golden/*.pt by running the REAL reference files
(/root/reference, through oracle/ref_shim.py) on seeded synthetic inputs and the deterministic weights
of oracle/pi0_oracle.make_pi0_weights.  Run in the authoring container only:

    python -m oracle.make_golden [tiny mid full verifier]

The fixtures are small (final actions + a few slices); the weights are regenerated from the seed by
whoever consumes them, so nothing large is committed.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import torch

from oracle import pi0_oracle as O
from oracle import ref_shim

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def pi0_truth(d, w, inp, K):
    """fp32 'truth' (SURVEY.md F10): the oracle graph with every bf16 rounding point removed, on the de-duplicated
    schedule (identical in exact arithmetic to the reference's batch layout; fp32 rounding differences ~1e-6)."""
    with O.truth_mode():
        return O.sample_actions_dedup(O.truth_weights(w), d, inp["image"], inp["tokens"], inp["masks"],
                                      inp["state"], inp["noise"], K)


def pi0_golden(name: str, R: int, K: int, seed: int = 0, truth: bool = False):
    d = getattr(O, name.upper())
    torch.manual_seed(0)
    t0 = time.time()
    model, _ = ref_shim.build_pi0(d.as_dict(), chunk_size=d.chunk_size, tokenizer_max_length=d.max_lang_len,
                                  num_steps=d.num_steps)
    w = O.make_pi0_weights(d, seed=seed)
    sd = model.state_dict()
    for k, v in w.items():
        sd[O.to_hf5_key(k)].copy_(v)
    inp = O.make_inputs(d, R, K, seed=seed)
    b = O.expand_to_batch(inp, K)
    N = R * K
    with torch.no_grad():
        img_emb = model.paligemma_with_expert.embed_image(b["image"][:1])
        embs, pad, att = model.embed_prefix([b["image"]], [torch.ones(N, dtype=torch.bool)], b["tokens"], b["masks"])
        M, _ = ref_shim.pi0_modules()
        mask = M.make_att_2d_masks(pad, att)
        pos = torch.cumsum(pad, dim=1) - 1
        _, cache = model.paligemma_with_expert.forward(attention_mask=mask, position_ids=pos, past_key_values=None,
                                                       inputs_embeds=[embs, None], use_cache=True, fill_kv_cache=True)
        v0 = model.denoise_step(b["state"], pad, cache, b["noise"].clone(), torch.tensor(1.0).expand(N))
        actions = model.sample_actions([b["image"]], [torch.ones(N, dtype=torch.bool)], b["tokens"], b["masks"],
                                       b["state"], noise=b["noise"].clone())
    L = d.layers - 1
    fix = dict(dims=d.as_dict(), R=R, K=K, seed=seed, actions=actions, v0=v0,
               image_emb_slice=img_emb[0, ::37, ::29].clone(),
               k0_slice=cache[0]["key_states"][::K, ::11, 0, ::7].clone(),
               vlast_slice=cache[L]["value_states"][::K, ::11, 0, ::7].clone(),
               lens=inp["lens"], torch_version=str(torch.__version__))
    if truth:
        t1 = time.time()
        fix["actions_truth"] = pi0_truth(d, w, inp, K)
        fix["err_ref_vs_truth"] = float((actions - fix["actions_truth"]).abs().max())
        print(f"  fp32 truth: {time.time() - t1:.1f}s  max|reference_bf16 - truth| = {fix['err_ref_vs_truth']:.3e}")
    OUT.mkdir(parents=True, exist_ok=True)
    torch.save(fix, OUT / f"pi0_{name}_R{R}K{K}.pt")
    print(f"pi0 {name} R={R} K={K}: {time.time() - t0:.1f}s  |actions-noise|max="
          f"{(actions - b['noise']).abs().max().item():.3f}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiny", "mid"]
    torch.set_num_threads(8)
    if "tiny" in which:
        pi0_golden("tiny", 2, 2)
    if "mid" in which:
        pi0_golden("mid", 2, 2)
    if "full" in which:
        pi0_golden("full", 2, 2, truth=True)
    if "full_r1k5" in which:  # BASELINE.json configs[1]
        pi0_golden("full", 1, 5, truth=True)
    if "full_r8k5" in which:  # BASELINE.json configs[2]
        pi0_golden("full", 8, 5, truth=True)
    if "mid_truth" in which:
        pi0_golden("mid", 2, 2, truth=True)
        pi0_golden("mid", 2, 3, truth=True)
    if "verifier" in which:
        from oracle import make_golden_verifier
        make_golden_verifier.main()
