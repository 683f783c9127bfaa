"""Generates tests/golden/exec_action.npz by EXECUTING the reference's own source (authoring container only):
BridgeSimplerAdapter.postprocess / postprocess_gripper (simpler.py:123-166, 211-220), BaseEnvAdapter.denormalize_bound
(base.py:20-31) - AST-extracted because the modules' import chains need absent packages - with the real
src/utils/geometry.py, and the vote lines run_simpler_eval_with_openpi.py:376-391 executed verbatim.

    python -m oracle.make_golden_exec
"""
from __future__ import annotations

import ast
import importlib.util
import json
import textwrap
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
ROOT = Path(__file__).resolve().parent.parent


def _method_source(path: Path, cls: str, name: str) -> str:
    src = path.read_text()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == name:
                    return textwrap.dedent(ast.get_source_segment(src, item))
    raise KeyError((cls, name))


def reference_adapter():
    """An object carrying the reference's unmodified postprocess / denormalize_bound / postprocess_gripper."""
    spec = importlib.util.spec_from_file_location("ref_geometry", REF / "INT-ACT/src/utils/geometry.py")
    geo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(geo)
    ns = {"np": np, "euler2axangle": geo.euler2axangle}
    simpler = REF / "INT-ACT/src/experiments/env_adapters/simpler.py"
    base = REF / "INT-ACT/src/experiments/env_adapters/base.py"
    exec(_method_source(simpler, "SimplerAdapter", "postprocess"), ns)
    exec(_method_source(simpler, "BridgeSimplerAdapter", "postprocess_gripper"), ns)
    exec(_method_source(base, "BaseEnvAdapter", "denormalize_bound"), ns)
    stats = json.loads((REF / "INT-ACT/config/dataset/bridge_statistics.json").read_text())

    class Adapter:
        action_normalization_type = "bound"
        dataset_statistics = stats
        postprocess = ns["postprocess"]
        postprocess_gripper = ns["postprocess_gripper"]
        denormalize_bound = ns["denormalize_bound"]

    return Adapter(), stats


def reference_vote(execution_action_histories_list, global_action_idx, K, num_past):
    """run_simpler_eval_with_openpi.py:372-391, the reference's lines executed verbatim."""
    lines = (REF / "CoVer_VLA/inference/experiments/robot/simpler/run_simpler_eval_with_openpi.py").read_text().splitlines()
    block = textwrap.dedent("\n".join(lines[371:391]))  # 'execute_action = ...' .. 'execute_action[-1] = float(np.sign(...))'
    assert block.lstrip().startswith("execute_action = execution_action_histories_list"), block[:80]

    class Cfg:
        policy_batch_inference_size = K

    ns = {"np": np, "execution_action_histories_list": execution_action_histories_list,
          "global_action_idx": global_action_idx, "num_past": num_past, "cfg": Cfg()}
    exec(block, ns)
    return ns["execute_action"], (ns["close_votes"], ns["open_votes"])


def reference_execution_action(adapter, actions, best_idx, K, step=0):
    N = actions.shape[0]
    hist = [adapter.postprocess(actions[n, step, :7].reshape(1, -1))[0][None, :] for n in range(N)]  # [1, 7] per candidate
    return reference_vote(hist, best_idx, K, 0)


def main():
    adapter, stats = reference_adapter()
    rng = np.random.default_rng(0)
    cases = []
    for R, K in [(8, 5), (3, 4), (1, 1), (2, 2)]:
        N = R * K
        a = rng.uniform(-1.3, 1.3, size=(N, 4, 32)).astype(np.float32)
        a[:, :, 6] = rng.uniform(0.0, 1.0, size=(N, 4)).astype(np.float32)
        if K == 2:
            a[0, 0, 6], a[1, 0, 6] = 0.9, 0.1  # a tie: the winner's own sign decides
        if K == 1:
            a[0, 0, 3:6] = 0.0
            a[0, 0, 3:6] = -(np.array(stats["action"]["p01"][3:6]) + np.array(stats["action"]["p99"][3:6])) / \
                (np.array(stats["action"]["p99"][3:6]) - np.array(stats["action"]["p01"][3:6]))  # ~identity rotation
        for idx in sorted({0, N - 1, N // 2, (N // 3)}):
            ex, votes = reference_execution_action(adapter, a, idx, K)
            cases.append((a, idx, K, ex, votes))
    out = {}
    for i, (a, idx, K, ex, votes) in enumerate(cases):
        out[f"a{i}"], out[f"idx{i}"], out[f"K{i}"], out[f"ex{i}"], out[f"v{i}"] = a, idx, K, ex, np.array(votes)
    out["n"] = len(cases)
    out["p01"] = np.array(stats["action"]["p01"][:6])
    out["p99"] = np.array(stats["action"]["p99"][:6])
    np.savez_compressed(ROOT / "tests/golden/exec_action.npz", **out)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
