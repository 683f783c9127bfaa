"""Generates tests/golden/format_traj.npz by EXECUTING the reference's own source (authoring container only):

  * process_inputs and convert_maniskill_with_bridge_adapter     CoVer_VLA/.../simpler/eval_utils.py:138-169, 172-221
    (AST-extracted, unmodified; only create_bridge_adapter_wrapper - which builds the adapter through an import chain
    that needs absent packages - is replaced by an object carrying the reference's unmodified methods)
  * SimplerAdapter.postprocess_verifier, BridgeSimplerAdapter.postprocess_gripper_verifier   INT-ACT/.../env_adapters/simpler.py:96-121, 222-226
  * BaseEnvAdapter.denormalize_bound                             INT-ACT/.../env_adapters/base.py:20-31
  * the -5 left padding to 10 steps                              bridge_verifier/ensemble_eval/efficient_ensemble_merged.py:378-390
    (the reference's lines executed verbatim)

    python -m oracle.make_golden_format
"""
from __future__ import annotations

import ast
import json
import textwrap
from pathlib import Path

import numpy as np
import torch

from oracle.make_golden_exec import REF, ROOT, _method_source

EVAL_UTILS = REF / "CoVer_VLA/inference/experiments/robot/simpler/eval_utils.py"
ENSEMBLE = REF / "bridge_verifier/ensemble_eval/efficient_ensemble_merged.py"


def _function_source(path: Path, name: str) -> str:
    src = path.read_text()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


def reference_process_inputs():
    """The reference's process_inputs, bound to an adapter that carries the reference's unmodified methods."""
    simpler = REF / "INT-ACT/src/experiments/env_adapters/simpler.py"
    base = REF / "INT-ACT/src/experiments/env_adapters/base.py"
    ns = {"np": np}
    exec(_method_source(simpler, "SimplerAdapter", "postprocess_verifier"), ns)
    exec(_method_source(simpler, "BridgeSimplerAdapter", "postprocess_gripper_verifier"), ns)
    exec(_method_source(base, "BaseEnvAdapter", "denormalize_bound"), ns)
    stats = json.loads((REF / "INT-ACT/config/dataset/bridge_statistics.json").read_text())

    class Adapter:
        action_normalization_type = "bound"
        dataset_statistics = stats
        postprocess_verifier = ns["postprocess_verifier"]
        postprocess_gripper_verifier = ns["postprocess_gripper_verifier"]
        denormalize_bound = ns["denormalize_bound"]

    mod = {"np": np, "create_bridge_adapter_wrapper": lambda temp: Adapter()}
    exec(_function_source(EVAL_UTILS, "convert_maniskill_with_bridge_adapter"), mod)
    exec(_function_source(EVAL_UTILS, "process_inputs"), mod)
    return mod["process_inputs"], stats


def reference_padding(all_action_histories):
    """efficient_ensemble_merged.py:378-390, the reference's lines executed verbatim -> f32 [N, 10, 7]."""
    lines = ENSEMBLE.read_text().splitlines()
    block = textwrap.dedent("\n".join(lines[378:390]))  # 'max_history_len = 10' .. 'action_histories_batch = torch.tensor(...)'
    assert block.lstrip().startswith("max_history_len = 10"), block[:60]
    assert "action_histories_batch = torch.tensor" in block.splitlines()[-1], block.splitlines()[-1]

    class Self:
        device = "cpu"

    ns = {"np": np, "torch": torch, "all_action_histories": all_action_histories, "self": Self()}
    exec(block, ns)
    return ns["action_histories_batch"]


def reference_trajectories(process_inputs, actions: np.ndarray, past, n_action_steps: int):
    """actions f32 [N, chunk, >=7], past = list of f64 [7] executed actions (the caller's action_history)."""
    N = actions.shape[0]

    class Cfg:
        pass

    Cfg.n_action_steps = n_action_steps
    # the deque select_action fills: n_action_steps tensors [batch, 7] (run_simpler_eval_with_openpi.py:322-334)
    queue = [torch.from_numpy(np.ascontiguousarray(actions[:, i, :7])) for i in range(n_action_steps)]
    hist = process_inputs(N, queue, verifier_action=True, action_history=list(past), cfg=Cfg())
    return reference_padding(hist).numpy()


def cases():
    rng = np.random.default_rng(1)
    for N, n_past in [(40, 0), (40, 2), (40, 6), (12, 9), (1, 1), (5, 5)]:
        a = rng.uniform(-1.4, 1.4, size=(N, 4, 32)).astype(np.float32)
        a[:, :, 6] = rng.uniform(0.0, 1.0, size=(N, 4)).astype(np.float32)
        a[0, 0, 6], a[0, 1, 6] = 0.5, np.nextafter(np.float32(0.5), np.float32(0))  # the gripper threshold itself
        a[0, 2, :6] = np.array([-1, 1, 0, -1, 1, 0], dtype=np.float32)               # the bounds themselves
        past = [np.concatenate([rng.uniform(-0.05, 0.05, size=6), [float(rng.integers(0, 2))]]) for _ in range(n_past)]
        yield a, past


def main():
    process_inputs, stats = reference_process_inputs()
    out = {"p01": np.array(stats["action"]["p01"][:6]), "p99": np.array(stats["action"]["p99"][:6])}
    n = 0
    for a, past in cases():
        out[f"a{n}"] = a
        out[f"past{n}"] = np.array(past, dtype=np.float64).reshape(len(past), 7)
        out[f"traj{n}"] = reference_trajectories(process_inputs, a, past, 4)
        n += 1
    out["n"] = n
    np.savez_compressed(ROOT / "tests/golden/format_traj.npz", **out)
    print("wrote", n, "cases")


if __name__ == "__main__":
    main()
