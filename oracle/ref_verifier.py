"""TEST INFRASTRUCTURE - builds the reference EfficientEnsembleMerged (real code from
/root/reference via oracle/ref_shim.py) without open_clip: components are constructed exactly as
efficient_ensemble_merged.py:96-184 does, weights come from oracle.verifier_oracle.make_verifier_weights,
and extract_shared_features is injected (the SigLIP2 trunk is third-party and absent - see
verifier_oracle.py).  Authoring-container only."""
from __future__ import annotations

from types import SimpleNamespace

import torch

from oracle import ref_shim
from oracle import verifier_oracle as V


def build_reference_ensemble(d: V.VerifierDims, w: dict, feature_fn):
    VM, EM = ref_shim.verifier_modules()
    ens = EM.EfficientEnsembleMerged.__new__(EM.EfficientEnsembleMerged)
    ens.device = "cpu"
    ens.use_transformer = d.traj_layers > 0
    ens.history_length = d.history
    ens.action_dim = d.action_dim
    ens.num_models = d.members
    ens.tokenizer = None
    ens.preprocess = None
    ens.siglip_model = SimpleNamespace(context_length=d.text_ctx)
    ens.trainable_models = []
    E = d.embed
    for m in range(d.members):
        b = f"verifier.{m}."

        def sub(prefix):
            return {k[len(prefix):]: v for k, v in w.items() if k.startswith(prefix)}

        text_aware = VM.TextAwareVisualExtraction(num_img_patches=d.n_patches, vision_dim=d.width)
        text_aware.load_state_dict(sub(b + "text_aware_visual_extraction."))
        vp = VM.AttentionPooling(input_dim=d.width, output_dim=E, num_heads=d.pool_heads, num_layers=d.pool_layers, num_readouts=1)
        vp.load_state_dict(sub(b + "vision_poolings."))
        tp = VM.AttentionPooling(input_dim=d.width, output_dim=E, num_heads=d.pool_heads, num_layers=d.pool_layers, num_readouts=1)
        tp.load_state_dict(sub(b + "text_pooling."))
        ip = torch.nn.Linear(2 * E, E)
        ip.load_state_dict(sub(b + "input_projection."))
        if d.traj_layers == 0:
            # the use_transformer = False components of efficient_ensemble_merged.py:161-183 (hidden width from the dims)
            ce = torch.nn.Sequential(torch.nn.Linear(d.history * d.action_dim, d.traj_ff), torch.nn.LayerNorm(d.traj_ff),
                                     torch.nn.ReLU(), torch.nn.Dropout(0.1), torch.nn.Linear(d.traj_ff, E))
            ce.load_state_dict(sub(b + "complex_action_encoder."))
            comps = {"text_aware_visual_extraction": text_aware.eval(), "vision_poolings": vp.eval(),
                     "text_pooling": tp.eval(), "input_projection": ip.eval(), "single_step_action_encoder": None,
                     "trajectory_encoder": None, "complex_action_encoder": ce.eval(), "action_padding_value": -5.0}
            ens.trainable_models.append(comps)
            continue
        ss = torch.nn.Linear(d.action_dim, E)
        ss.load_state_dict(sub(b + "single_step_action_encoder."))
        layer = torch.nn.TransformerEncoderLayer(d_model=E, nhead=d.pool_heads, dim_feedforward=d.traj_ff, batch_first=False, dropout=0.1)
        te = torch.nn.TransformerEncoder(layer, num_layers=d.traj_layers)
        te.load_state_dict(sub(b + "trajectory_encoder."))
        comps = {"text_aware_visual_extraction": text_aware.eval(), "vision_poolings": vp.eval(), "text_pooling": tp.eval(),
                 "input_projection": ip.eval(), "single_step_action_encoder": ss.eval(), "trajectory_encoder": te.eval(),
                 "complex_action_encoder": None, "action_padding_value": -5.0}
        ens.trainable_models.append(comps)
    ens.extract_shared_features = feature_fn
    return ens
