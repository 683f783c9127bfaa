"""PI0Policy host surface on the CPU: prepare_language with a stand-in tokenizer (what the tokenizer
is CALLED with and what is done with its output), select_action's queue semantics and the input / output normalisation -
each against the reference's own code executed live where /root/reference is present (modeling_pi0.py:263-307, 388-409;
normalize.py:116-254)."""
import ast
import textwrap
import types
from pathlib import Path

import pytest
import torch

from cover_vla_b200.pi0.configuration_pi0 import PI0Config
from cover_vla_b200.pi0.modeling_pi0 import OBS_ROBOT, PI0Policy

REF = Path("/root/reference/lerobot_custom/lerobot/common/policies/pi0/modeling_pi0.py")


class FakeTokenizer:
    """Deterministic word-hash tokenizer with the call signature of a HF tokenizer; remembers how it was called."""

    def __init__(self):
        self.calls = []

    def __call__(self, texts, **kw):
        self.calls.append((list(texts), dict(kw)))
        L = kw["max_length"]
        ids = torch.zeros(len(texts), L, dtype=torch.int64)
        mask = torch.zeros(len(texts), L, dtype=torch.int64)
        for i, t in enumerate(texts):
            toks = [2] + [3 + (sum(map(ord, w)) % 997) for w in t.replace("\n", " \n").split(" ") if w]
            toks = toks[:L] if kw.get("truncation") else toks
            assert len(toks) <= L and kw.get("padding") == "max_length" and kw.get("padding_side") == "right"
            ids[i, :len(toks)] = torch.tensor(toks)
            mask[i, :len(toks)] = 1
        return {"input_ids": ids, "attention_mask": mask}


def _policy(tok, max_len=16):
    cfg = PI0Config.bridge()
    cfg.tokenizer_max_length = max_len
    engine = types.SimpleNamespace(cfg=types.SimpleNamespace(num_cameras=1), device="cpu")
    return PI0Policy(cfg, engine=engine, language_tokenizer=tok)


TASKS = ["put the spoon on the towel", "place spoon onto the towel\n", "move the carrot to the plate and then stop moving at all "
         "and wait for the next instruction to arrive"]


def test_tokenizer_branch_calls_and_cache():
    tok = FakeTokenizer()
    pol = _policy(tok)
    batch = {OBS_ROBOT: torch.zeros(3, 7), "task": list(TASKS)}
    ids, masks = pol.prepare_language(batch)
    texts, kw = tok.calls[0]
    assert texts == [TASKS[0] + "\n", TASKS[1], TASKS[2] + "\n"]              # the prompt has to end with a new line (:394-395)
    assert kw == dict(padding="max_length", padding_side="right", max_length=16, return_tensors="pt", truncation=True)
    assert ids.dtype == torch.int64 and masks.dtype == torch.bool and ids.shape == masks.shape == (3, 16)
    assert bool(masks[2].all())                                               # the long prompt is truncated to max_length
    assert pol.model.lang_len_hint == int(masks.sum(1).max()) == 16          # host-known bound on the valid length
    # per-task prompt cache: the same prompt set is not tokenised again and returns the same tensors
    ids2, masks2 = pol.prepare_language(batch)
    assert len(tok.calls) == 1 and torch.equal(ids, ids2) and torch.equal(masks, masks2)
    # another prompt set (the instruction swap of run_simpler_eval_with_openpi.py:409) is tokenised, with its own bound
    pol.prepare_language({OBS_ROBOT: torch.zeros(2, 7), "task": TASKS[:2]})
    assert len(tok.calls) == 2 and pol.model.lang_len_hint < 16
    # pre-tokenised prompts must not inherit that bound
    pol.prepare_language({OBS_ROBOT: torch.zeros(1, 7), "lang_tokens": ids[:1], "lang_masks": masks[:1]})
    assert pol.model.lang_len_hint is None
    with pytest.raises(RuntimeError):
        _policy(None).prepare_language(batch)


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the authoring container")
def test_tokenizer_branch_matches_the_reference_method_live():
    src = REF.read_text()
    fn = None
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == "PI0Policy":
            fn = next(i for i in node.body if isinstance(i, ast.FunctionDef) and i.name == "prepare_language")
    code = textwrap.dedent(ast.get_source_segment(src, fn))
    code = code[code.index("def prepare_language"):]                          # drop the @torch.compiler.disable decorator
    ns = {"torch": torch, "Tensor": torch.Tensor, "OBS_ROBOT": OBS_ROBOT}
    exec(code, ns)
    ref_tok, our_tok = FakeTokenizer(), FakeTokenizer()
    ref_self = types.SimpleNamespace(language_tokenizer=ref_tok, config=types.SimpleNamespace(tokenizer_max_length=16))
    batch = {OBS_ROBOT: torch.zeros(3, 7), "task": list(TASKS)}
    ref_ids, ref_masks = ns["prepare_language"](ref_self, batch)
    ids, masks = _policy(our_tok).prepare_language(batch)
    assert our_tok.calls == ref_tok.calls                                      # same texts, same keyword arguments
    assert torch.equal(ids, ref_ids) and torch.equal(masks, ref_masks) and masks.dtype == ref_masks.dtype


# ---------------------------------------------------------------------------------------------------------------------
# select_action: queue / refill / slicing / un-pad / un-normalise semantics against the reference's own
# method executed live, with the sampler replaced by the same deterministic stand-in on both sides
# ---------------------------------------------------------------------------------------------------------------------
def _sampler(log):
    def sample_actions(images, img_masks, lang_tokens, lang_masks, state, noise=None, noise_std=1.0):
        log.append((len(images), tuple(lang_tokens.shape), tuple(state.shape), noise_std))
        g = torch.Generator().manual_seed(len(log))
        return torch.randn(state.shape[0], 4, 32, generator=g)
    return sample_actions


def _batch(n):
    return {"observation.images.top": torch.rand(n, 3, 224, 224) * 2 - 1, OBS_ROBOT: torch.randn(n, 7),
            "lang_tokens": torch.randint(3, 100, (n, 16)), "lang_masks": torch.ones(n, 16, dtype=torch.bool)}


def _bridge_policy(n_action_steps, stats=None, mapping=None):
    from cover_vla_b200.pi0.configuration_pi0 import PolicyFeature
    cfg = PI0Config.bridge()
    cfg.tokenizer_max_length, cfg.n_action_steps = 16, n_action_steps
    cfg.input_features = {"observation.images.top": PolicyFeature("VISUAL", (3, 224, 224)), OBS_ROBOT: PolicyFeature("STATE", (7,))}
    cfg.output_features = {"action": PolicyFeature("ACTION", (7,))}
    if mapping:
        cfg.normalization_mapping = mapping
    engine = types.SimpleNamespace(cfg=types.SimpleNamespace(num_cameras=1), device="cpu")
    return PI0Policy(cfg, engine=engine, dataset_stats=stats)


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the authoring container")
@pytest.mark.parametrize("n_action_steps,normalise", [(4, False), (2, False), (4, True)])
def test_select_action_queue_semantics_match_the_reference_method_live(n_action_steps, normalise):
    from collections import deque
    src = REF.read_text()
    fn = None
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.ClassDef) and node.name == "PI0Policy":
            fn = next(i for i in node.body if isinstance(i, ast.FunctionDef) and i.name == "select_action")
    code = textwrap.dedent(ast.get_source_segment(src, fn))
    code = code[code.index("def select_action"):]                             # drop the @torch.no_grad decorator
    ns = {"torch": torch, "Tensor": torch.Tensor, "OBS_ROBOT": OBS_ROBOT}
    exec(code, ns)
    stats = mapping = None
    if normalise:
        stats = {"action": {"mean": torch.linspace(-1, 1, 7), "std": torch.linspace(0.5, 2, 7)}}
        mapping = {"VISUAL": "IDENTITY", "STATE": "IDENTITY", "ACTION": "MEAN_STD"}
    ours, helper = _bridge_policy(n_action_steps, stats, mapping), _bridge_policy(n_action_steps, stats, mapping)
    log_o, log_r = [], []
    ours.model.sample_actions = _sampler(log_o)
    # the reference method on an object that borrows the mirror's prepare_* / normalise helpers, so both sides hand the
    # sampler the same tensors and the comparison isolates select_action itself
    ref = types.SimpleNamespace(config=helper.config, eval=lambda: None, normalize_inputs=helper.normalize_inputs,
                                unnormalize_outputs=helper.unnormalize_outputs, prepare_images=helper.prepare_images,
                                prepare_state=helper.prepare_state, prepare_language=helper.prepare_language,
                                model=types.SimpleNamespace(sample_actions=_sampler(log_r)),
                                _action_queue=deque([], maxlen=n_action_steps))
    N = 6
    for it in range(3):
        b = _batch(N)
        q_ref = ns["select_action"](ref, dict(b), noise_std=0.7)
        q_our = ours.select_action(dict(b), noise_std=0.7)
        assert isinstance(q_our, deque) and q_our.maxlen == q_ref.maxlen == n_action_steps       # the deque itself (:307)
        assert len(q_our) == len(q_ref) == n_action_steps
        assert all(torch.equal(a, r) and a.shape == (N, 7) for a, r in zip(q_our, q_ref))
        assert log_o == log_r
        if it == 0:  # a second call on a non-empty queue does not sample again (:277)
            ns["select_action"](ref, dict(b))
            ours.select_action(dict(b))
            assert len(log_o) == len(log_r) == 1
        # the caller copies and clears (run_simpler_eval_with_openpi.py:324-326); the next call refills
        q_ref.clear()
        q_our.clear()
    assert len(log_o) == 3


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the authoring container")
@pytest.mark.parametrize("mode", ["MEAN_STD", "MIN_MAX", "IDENTITY"])
def test_normalisation_matches_the_reference_modules_live(mode):
    """normalize_inputs / unnormalize_outputs of the mirror (pi0/modeling_pi0.py _Normalize) against the reference's
    Normalize / Unnormalize nn.Modules (lerobot/common/policies/normalize.py:116-254), bit for bit."""
    from oracle import ref_shim
    ref_shim.install()
    from lerobot.common.policies.normalize import Normalize, Unnormalize
    from lerobot.configs.types import FeatureType, NormalizationMode
    from lerobot.configs.types import PolicyFeature as RefFeature
    from cover_vla_b200.pi0.configuration_pi0 import PolicyFeature
    from cover_vla_b200.pi0.modeling_pi0 import _Normalize
    g = torch.Generator().manual_seed(3)
    stats = {OBS_ROBOT: {"mean": torch.randn(7, generator=g), "std": torch.rand(7, generator=g) + 0.1,
                         "min": -torch.rand(7, generator=g) - 0.5, "max": torch.rand(7, generator=g) + 0.5},
             "action": {"mean": torch.randn(7, generator=g), "std": torch.rand(7, generator=g) + 0.1,
                        "min": -torch.rand(7, generator=g) - 0.5, "max": torch.rand(7, generator=g) + 0.5}}
    ref_in = {OBS_ROBOT: RefFeature(type=FeatureType.STATE, shape=(7,))}
    ref_out = {"action": RefFeature(type=FeatureType.ACTION, shape=(7,))}
    ref_map = {FeatureType.STATE: NormalizationMode[mode], FeatureType.ACTION: NormalizationMode[mode],
               FeatureType.VISUAL: NormalizationMode.IDENTITY}
    our_map = {"STATE": mode, "ACTION": mode, "VISUAL": "IDENTITY"}
    ours_n = _Normalize({OBS_ROBOT: PolicyFeature("STATE", (7,))}, our_map, stats, False)
    ours_u = _Normalize({"action": PolicyFeature("ACTION", (7,))}, our_map, stats, True)
    ref_n, ref_u = Normalize(ref_in, ref_map, stats), Unnormalize(ref_out, ref_map, stats)
    x = {OBS_ROBOT: torch.randn(5, 7, generator=g), "other": torch.ones(2)}
    a = {"action": torch.randn(5, 4, 7, generator=g)}
    assert torch.equal(ours_n(x)[OBS_ROBOT], ref_n(x)[OBS_ROBOT]) and torch.equal(ours_n(x)["other"], x["other"])
    assert torch.equal(ours_u(a)["action"], ref_u(a)["action"])
    if mode != "IDENTITY":  # without statistics the mirror refuses (the reference asserts on its infinite placeholder buffers)
        with pytest.raises(ValueError):
            _Normalize({OBS_ROBOT: PolicyFeature("STATE", (7,))}, our_map, None, False)(x)


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the authoring container")
def test_image_and_state_glue_matches_the_reference_functions_live():
    """resize_with_pad / pad_vector (modeling_pi0.py:131-165) incl. the mirror's shortcut for frames that already have the
    target size (the reference interpolates them at scale 1, which returns the same bits)."""
    from oracle import ref_shim
    ref_shim.install()
    import lerobot.common.policies.pi0.modeling_pi0 as R
    from cover_vla_b200.pi0 import modeling_pi0 as M
    g = torch.Generator().manual_seed(9)
    for shape in [(2, 3, 224, 224), (1, 3, 480, 640), (2, 3, 100, 300), (1, 3, 256, 200)]:
        img = torch.rand(*shape, generator=g) * 2 - 1
        for pad in (0, -1):
            assert torch.equal(M.resize_with_pad(img, 224, 224, pad_value=pad), R.resize_with_pad(img, 224, 224, pad_value=pad))
    for shape in [(5, 7), (5, 32), (2, 4, 7)]:
        v = torch.randn(*shape, generator=g)
        assert torch.equal(M.pad_vector(v, 32), R.pad_vector(v, 32))
