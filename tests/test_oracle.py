"""The oracle (oracle/vap_oracle.py) against the reference's own outputs.

The golden arrays were produced by the UNMODIFIED reference on CPU fp32
(oracle/make_golden.py). The oracle uses the same ATen ops in the same order, so
on the same torch build it is bit-identical; across builds we allow 2e-6.
"""
import os

import pytest
import torch

from conftest import CASE_NAMES, golden_inputs, load_golden
from oracle import ref_import, synth
from oracle import vap_oracle as O

TOL = 2e-6


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_reference_golden(name):
    if name == "lstm1_turns_T1000" and os.environ.get("VAP_FAST_TESTS"):
        pytest.skip("fast mode")
    recipe, g = load_golden(name)
    sd, wav = golden_inputs(recipe, g)
    stages = {}
    fwd = O.forward(sd, wav, stages=stages)
    assert (fwd["logits"] - g["logits"]).abs().max() <= 5e-5  # logits are O(1..10)
    assert (fwd["vad"] - g["vad_logits"]).abs().max() <= 5e-5
    out = O.probs(sd, wav)
    assert list(out.keys()) == ["probs", "vad", "p_now", "p_future", "H", "loss"]
    for k in ["probs", "vad", "p_now", "p_future"]:
        if k in g:
            assert out[k].shape == g[k].shape
            assert (out[k] - g[k]).abs().max() <= TOL, k
    assert (out["H"] - g["H"]).abs().max() <= 2e-5
    assert out["loss"].shape == g["loss"].shape
    assert (out["loss"] - g["loss"]).abs().max() <= 5e-5
    # decisions are bit-exact
    assert torch.equal(fwd["logits"].argmax(-1), g["logits"].argmax(-1))
    assert torch.equal(out["vad"] >= 0.5, g["vad"] >= 0.5)
    if recipe.get("stages"):
        m = {"stage_conv": "conv_1", "stage_ar": "ar_1", "stage_enc": "enc_1", "stage_ch": "ch_1",
             "stage_comb": "comb"}
        for gk, sk in m.items():
            assert (stages[sk][:1] - g[gk]).abs().max() <= 2e-5, gk
        for l in range(3):
            for c in (1, 2):
                assert (stages[f"ar{l}_x{c}"][:1] - g[f"stage_ar{l}_x{c}"]).abs().max() <= 5e-5


def test_oracle_session_stitching_matches_reference_golden():
    recipe, g = load_golden("session_45s")
    sd = synth.make_state_dict(recipe["seed"], recipe["ar_mode"], recipe["ar_layers"], recipe["gain"])
    for tag in ("a", "b"):
        n = int(g[f"{tag}_n_samples"])
        wav = synth.make_waveform(1, n, recipe["wav_seed"], recipe["kind"])
        out = O.step_extraction(sd, wav)
        for k in ["vad", "p_now", "p_future"]:
            assert out[k].shape == g[f"{tag}_{k}"].shape
            assert (out[k] - g[f"{tag}_{k}"]).abs().max() <= TOL
        assert (out["H"] - g[f"{tag}_H"]).abs().max() <= 2e-5
        assert (out["loss"] - g[f"{tag}_loss"]).abs().max() <= 5e-5
        assert torch.equal(out["probs"].argmax(-1).to(torch.uint8), g[f"{tag}_probs_argmax"])


def test_recurrence_equations_match_aten():
    for mode, layers in [("LSTM", 1), ("GRU", 1), ("LSTM", 2)]:
        sd = synth.make_state_dict(11, mode, layers)
        z = torch.randn(2, 40, 256, generator=torch.Generator().manual_seed(0))
        assert (O.ar_net(sd, z) - O.ar_net_loop(sd, z)).abs().max() < 1e-5
        assert O.ar_kind(sd) == (mode, layers)


def test_frame_count_chain():
    # SURVEY.md F10b probe table
    assert O.n_frames(37392)[-2:] == [233, 117]
    assert O.n_frames(160000)[-2:] == [1000, 500]
    assert O.n_frames(320000) == [64000, 16000, 8000, 4000, 2000, 1000]
    assert O.n_frames(400000)[-2:] == [2500, 1250]
    assert O.n_frames(32159)[-1] == 101 and O.n_frames(32158)[-1] == 100


def test_probs_raises_for_100_frames_or_fewer():
    # reference vap/model.py:220-224 via objective.py:53 (unfold), SURVEY.md F6
    sd = synth.make_state_dict(0)
    wav = synth.make_waveform(1, 32000, 0)
    with pytest.raises(RuntimeError):
        O.probs(sd, wav)


def test_codebook_and_alibi():
    cv = O.code_vectors(8)
    assert cv[1].tolist() == [1, 0, 0, 0, 0, 0, 0, 0]
    assert cv[16].tolist() == [0, 0, 0, 0, 1, 0, 0, 0]
    assert synth.alibi_slopes(4) == [0.25, 0.0625, 0.015625, 0.00390625]
    m = O.alibi_mask(torch.tensor(synth.alibi_slopes(4)), 5)
    assert m[0, 0, 2, 1] == 1.0 + 0.25 and m[0, 0, 1, 2] == float("-inf")


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_oracle_bit_identical_to_imported_reference():
    sd = synth.make_state_dict(21, "LSTM", 1, 2.0)
    model = ref_import.build_reference(sd)
    wav = synth.make_waveform(1, 36000, 5, "turns")
    ref = model.probs(wav)
    mine = O.probs(sd, wav)
    for k in ref:
        assert torch.equal(ref[k], mine[k]), k


# --------------------------------------------------------------------------- #
# resampler oracle (input path)                           #
# --------------------------------------------------------------------------- #
def test_resample_oracle_matches_torchaudio_golden():
    import numpy as np

    from conftest import GOLDEN_DIR
    from oracle import resample_oracle as R

    g = np.load(f"{GOLDEN_DIR}/resample_example_24k_16k.npz")
    assert np.abs(R.resample(g["pcm_24k"], 24000, 16000) - g["y_24k"]).max() <= 1e-6
    for rate in (48000, 44100, 8000):
        y = R.resample(g[f"x_{rate}"], rate, 16000)
        assert y.shape == g[f"y_{rate}"].shape
        assert np.abs(y - g[f"y_{rate}"]).max() <= 1e-6


def test_resample_bank_is_torchaudios_and_oracle_matches_torchaudio_live():
    import math

    import numpy as np

    from oracle import resample_oracle as R
    from voiceactivityprojection_b200.audio import sinc_resample_bank

    taf = pytest.importorskip("torchaudio.functional")
    rng = np.random.default_rng(0)
    for rate, n in ((24000, 1000), (24000, 1), (48000, 3333), (44100, 900), (8000, 555), (32000, 77), (16000 * 3, 10)):
        bank, width, orig, new = sinc_resample_bank(rate, 16000)
        kb, w = taf.functional._get_sinc_resample_kernel(rate, 16000, math.gcd(rate, 16000), dtype=torch.float32)
        assert w == width and torch.equal(kb[:, 0], bank)
        assert np.abs(R.bank(rate, 16000)[0] - bank.numpy()).max() <= 2e-7
        x = (rng.random((2, n), dtype=np.float32) * 2 - 1)
        ref = taf.resample(torch.from_numpy(x), rate, 16000).numpy()
        got = R.resample(x, rate, 16000)
        assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-6


def test_vad_filter_oracle_matches_reference_golden():
    import numpy as np

    from conftest import GOLDEN_DIR

    g = np.load(f"{GOLDEN_DIR}/vad_filter.npz")
    names = sorted(k[:-3] for k in g.files if k.endswith("_in"))
    assert len(names) == 5
    for n in names:
        fill, omit = g[n + "_par"]
        got = O.vad_filter(torch.from_numpy(g[n + "_in"]), float(fill), float(omit))
        assert torch.equal(got, torch.from_numpy(g[n + "_out"])), n


def test_zero_shot_oracle_matches_reference_golden():
    """oracle/zero_shot_oracle.py against the unmodified reference's ZeroShot (oracle/make_golden_zeroshot.py):
    class subsets bit-exact, marginals within 1e-6 (float32 sums of <= 56 probabilities in a different order)."""
    import numpy as np

    from conftest import GOLDEN_DIR
    from oracle import zero_shot_oracle as Z

    g = np.load(f"{GOLDEN_DIR}/zero_shot.npz")
    for k, v in Z.subsets().items():
        assert np.array_equal(v, g[k]), k
    for name in ("flat", "peaked", "odd"):
        r = Z.get_probs(g[name + "_logits"], g[name + "_va"])
        for k in ("p", "p_bc", "p_sil", "p_act"):
            assert r[k].shape == g[f"{name}_{k}"].shape
            assert np.abs(r[k] - g[f"{name}_{k}"]).max() <= 1e-6, (name, k)
        # from probabilities instead of logits: same numbers
        e = np.exp(g[name + "_logits"] - g[name + "_logits"].max(-1, keepdims=True))
        r2 = Z.get_probs(e / e.sum(-1, keepdims=True), g[name + "_va"], is_probs=True)
        assert np.array_equal(r2["p"], r["p"])


def test_oracle_attention_maps_match_reference_golden():
    """forward(attention=True): oracle maps bit-identical to the unmodified reference's
    (oracle/make_golden_attention.py), shapes (B, 2, layers, H, T, T), and unchanged logits."""
    recipe, arrays = load_golden("attention_maps_T70")
    sd, wav = golden_inputs(recipe, arrays)
    with torch.no_grad():
        out = O.forward(sd, wav, attention=True)
        plain = O.forward(sd, wav)
    assert out["self_attn"].shape == (1, 2, 1, 4, 70, 70) and out["cross_attn"].shape == (1, 2, 3, 4, 70, 70)
    for k in ("logits", "vad", "self_attn", "cross_attn", "cross_self_attn"):
        assert torch.equal(out[k], arrays[k]), k
    assert torch.equal(out["logits"], plain["logits"]) and set(plain) == {"logits", "vad"}
