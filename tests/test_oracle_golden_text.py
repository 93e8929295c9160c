"""End-to-end pin of the oracle: the reference's golden generations
(crabml-llama2/src/llama2.rs:673-703) on its own GGUF fixtures, greedy decode, F32 KV cache."""
import os

import numpy as np
import pytest

from oracle.llama_replay import GGUFModel, Llama2Runner, LlamaTokenizer, decode_text, load_weights
from oracle.tensor_ref import OracleDevice, OracleTensor

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PROMPT = "Lily is a cute cat, "
PROMPT_IDS = [1, 365, 2354, 338, 263, 274, 1082, 6635, 29892, 29871]

CASES = [
    ("tinyllamas-stories-15m-q8_0.gguf", "3 years old. She likes to play with her",        # llama2.rs:683-685
     [29941, 2440, 2030, 29889, 2296, 4188, 267, 304, 1708, 411, 902]),
    ("tinyllamas-stories-15m-q4_0.gguf", "3 year old Lily. She likes to play",             # llama2.rs:699-701
     [29941, 1629, 2030, 365, 2354, 29889, 2296, 4188, 267, 304, 1708]),
]


def run_oracle(path, steps=11, f16_kv=False, debug=False):
    gm = GGUFModel(path)
    dev = OracleDevice(debug_named_tensors=debug)
    w = load_weights(gm, OracleTensor, dev)
    tok = LlamaTokenizer(gm.tokens, gm.scores, gm.bos, gm.eos)
    ids = tok.encode(PROMPT, True, False)
    r = Llama2Runner(OracleTensor, gm.conf, w, dev, 200, use_f16_kv_cache=f16_kv)
    logits = []
    pos, _, t0 = r.prefill(ids)
    logits.append(r.logits.copy())
    out = []
    for t in r.generate(pos, t0, steps, eos=gm.eos):
        out.append(t)
        logits.append(r.logits.copy())
    return gm, tok, ids, out, logits, dev


@pytest.mark.parametrize("fname,text,ids", CASES)
def test_golden_generation(fixture_path, fname, text, ids):
    gm, tok, prompt_ids, out, _, _ = run_oracle(fixture_path(fname))
    assert gm.conf.rope_dim == 48 and gm.conf.head_size() == 48          # llama2.rs:679-680
    assert prompt_ids == PROMPT_IDS
    assert out == ids
    assert decode_text(tok, out) == text


def test_golden_generation_f16_kvcache(fixture_path):
    # llama2.rs:706-719: same text with the F16 KV cache
    _, tok, _, out, _, _ = run_oracle(fixture_path("tinyllamas-stories-15m-q8_0.gguf"), f16_kv=True)
    assert decode_text(tok, out) == "3 years old. She likes to play with her"


@pytest.mark.parametrize("fname", [c[0] for c in CASES])
def test_committed_logits_fixture(fixture_path, fname):
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from the oracle on the
    reference's GGUF files; this guards the oracle (and the fixtures) against drift."""
    npz = os.path.join(GOLDEN, fname.replace(".gguf", "_logits.npz"))
    if not os.path.exists(npz):
        pytest.skip("golden npz not generated")
    g = np.load(npz)
    _, _, _, out, logits, _ = run_oracle(fixture_path(fname))
    assert out == g["generated_ids"].tolist()
    np.testing.assert_array_equal(np.stack(logits)[g["steps"]], g["logits"])
