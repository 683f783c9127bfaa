"""TEST INFRASTRUCTURE - generates tests/golden/verifier_*.pt by running the REAL reference
EfficientEnsembleMerged.compute_max_similarity_scores_batch (bridge_verifier/ensemble_eval/
efficient_ensemble_merged.py:309-454, via oracle/ref_shim.py + oracle/ref_verifier.py) on seeded synthetic
inputs.  The SigLIP2 trunk is third-party and absent from /root/reference, so `extract_shared_features`
is injected with the oracle's trunk restatement (verifier_oracle.extract_features): these fixtures pin the
HEADS, fusion and selection rule (reference code), not the trunk (unpinned upstream).

    python -m oracle.make_golden_verifier          (authoring container only)
"""
from __future__ import annotations

from pathlib import Path

import torch

from oracle import ref_verifier
from oracle import verifier_oracle as V

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def run_reference(d: V.VerifierDims, w: dict, inp: dict, group_size: int):
    """The reference object, called the way run_simpler_eval_with_openpi.py:355-363 calls it."""
    patch, text = V.extract_features(w, d, inp["image"], inp["tokens"])

    def feature_fn(img_tensor, text_tokens):
        # efficient_ensemble_merged.py:188-192 returns (patch_features, text_features) for one pair
        return patch, text

    ens = ref_verifier.build_reference_ensemble(d, w, feature_fn)
    # pre-tokenised instructions + a preprocess stub: the branch of :330-347 only needs one image tensor
    ens.preprocess = lambda im: inp["image"][0]
    N = len(inp["histories"])
    captured = {}
    orig = ens.get_embeddings_from_model_batch

    def spy(model_idx, patch_features, text_features, action_histories):
        it, act = orig(model_idx, patch_features, text_features, action_histories)
        captured.setdefault("it", []).append(it[0].clone())
        captured.setdefault("act", []).append(act.clone())
        return it, act

    ens.get_embeddings_from_model_batch = spy
    with torch.no_grad():
        ms, mi, mh, gi = ens.compute_max_similarity_scores_batch(
            [inp["image"][0]] * N, [inp["tokens"][0]] * N, inp["histories"], cfg_repeat_language_instructions=group_size)
    return dict(max_score=float(ms), global_idx=int(gi), it=torch.stack(captured["it"]), act=torch.stack(captured["act"]),
                patch=patch, text=text)


def make(name: str, R: int, K: int, seed: int):
    d = getattr(V, name)
    w = V.make_verifier_weights(d, seed=0)
    inp = V.make_inputs(d, R * K, seed=seed)
    ref = run_reference(d, w, inp, K)
    # scores the reference consumes = row 0 of fused_it @ fused_act.T (:414-425)
    fit = ref["it"].mean(0, keepdim=True)
    fit = fit / fit.norm(dim=-1, keepdim=True)
    fact = ref["act"].mean(0)
    fact = fact / fact.norm(dim=-1, keepdim=True)
    scores = (fit @ fact.T)[0]
    fix = dict(name=name, R=R, K=K, seed=seed, max_score=ref["max_score"], global_idx=ref["global_idx"], scores=scores,
               it_emb=ref["it"], act_emb_slice=ref["act"][:, ::3, ::17].clone(), torch_version=str(torch.__version__),
               # trunk restatement outputs (strided slices): pins the CUDA trunk at sizes the GPU tests do not recompute
               patch_slice=ref["patch"][0, ::7, ::13].clone(), text_slice=ref["text"][0, ::3, ::13].clone())
    OUT.mkdir(parents=True, exist_ok=True)
    torch.save(fix, OUT / f"verifier_{name.lower()}_R{R}K{K}.pt")
    print(f"verifier {name} R={R} K={K}: max_score {ref['max_score']:.6f} idx {ref['global_idx']}")


def main(full: bool = False):
    torch.set_num_threads(8)
    if full:  # BASELINE.json configs[0] / configs[2]: full-size trunk + heads, 8 rephrases x 5 samples
        make("VFULL", 8, 5, seed=2)
        return
    make("VTINY", 4, 3, seed=1)
    make("VMID", 8, 5, seed=2)
    make("VMID", 1, 1, seed=3)
    # use_transformer = False checkpoints (MLP action encoder, efficient_ensemble_merged.py:161-171, 241-245)
    make("VTINY_MLP", 4, 3, seed=1)
    make("VMID_MLP", 8, 5, seed=2)


if __name__ == "__main__":
    import sys
    main(full="full" in sys.argv[1:])
