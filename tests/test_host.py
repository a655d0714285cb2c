"""CPU tests of the host side: C-ABI surface, session stitching, wav loading,
config/argparse mirror, JSON contract. No kernel is launched here."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_in_header():
    from voiceactivityprojection_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "vapb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(vapb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/vapb.h but not exported by libvapb.so"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert b"sm_100a" in lib.vapb_build_info()


def test_frames_chain_matches_oracle():
    from oracle import vap_oracle as O
    from voiceactivityprojection_b200 import _lib

    for n in [32159, 37392, 160000, 320000, 400000, 9600000, 33333]:
        assert _lib.frames(n) == tuple(O.n_frames(n))[-2:]
    with pytest.raises(_lib.VapbError):
        _lib.frames(3)


def test_no_cpu_fallback():
    from oracle import synth
    from voiceactivityprojection_b200 import VapConfig, VapGPT

    m = VapGPT(VapConfig())
    m.load_state_dict(synth.make_state_dict(0))
    with pytest.raises(RuntimeError):
        m.probs(torch.zeros(1, 2, 40000))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 2, 40000))
    with pytest.raises(NotImplementedError):
        m.train()
    lg, tgt = torch.tensor([[0.0, 2.0]]), torch.tensor([[1.0, 0.0]])
    assert torch.allclose(m.vad_loss(lg, tgt), torch.nn.functional.binary_cross_entropy_with_logits(lg, tgt))


class _FakeModel:
    """Cheap stand-in with the facade's probs() contract: frame features are
    deterministic functions of the window's samples, so stitching is checkable."""
    sample_rate, frame_hz = 16000, 50

    def probs(self, w, **kw):
        B, _, S = w.shape
        T = S // 320
        f = w[..., : T * 320].reshape(B, 2, T, 320)
        a = f.mean(-1).transpose(1, 2)                      # (B, T, 2)
        pos = torch.arange(T, dtype=w.dtype, device=w.device)[None, :, None]  # position inside the window matters
        return {"probs": (a.sum(-1, keepdim=True) + pos).repeat(1, 1, 4), "vad": a + pos, "p_now": a - pos,
                "p_future": a * 2 + pos, "H": a.sum(-1) + pos[..., 0], "loss": a[:, : T - 100, 0]}


def _reference_loop(wav, model, context_time=20, step_time=5):
    """Restatement of run.py:23-131 (one forward per window, concatenation)."""
    n = wav.shape[-1]
    duration = round(n / model.sample_rate, 2)
    cs, ss = int((context_time + step_time) * model.sample_rate), int(step_time * model.sample_rate)
    sf = int(step_time * model.frame_hz)
    folds = wav.unfold(dimension=-1, size=cs, step=ss).permute(2, 0, 1, 3)
    out = model.probs(folds[0])
    keys = ["vad", "p_now", "p_future", "probs", "H"]
    for w in folds[1:]:
        o = model.probs(w)
        for k in keys:
            out[k] = torch.cat([out[k], o[k][:, -sf:]], dim=1)
    if round(duration * model.frame_hz) != out["p_now"].shape[1]:
        om = round(duration * model.frame_hz) - out["p_now"].shape[1]
        o = model.probs(wav[..., -cs:])
        for k in keys:
            out[k] = torch.cat([out[k], o[k][:, -om:]], dim=1)
    return out


@pytest.mark.parametrize("B,seconds,max_batch", [(1, 45.0, 64), (2, 61.3, 3), (1, 25.0, 8), (3, 40.0, 4), (1, 172.7, 16)])
def test_batched_session_stitching_equals_reference_loop(B, seconds, max_batch):
    from voiceactivityprojection_b200.session import step_extraction

    g = torch.Generator().manual_seed(int(seconds * 10) + B)
    wav = torch.randn(B, 2, int(seconds * 16000), generator=g)
    m = _FakeModel()
    got = step_extraction(wav, m, device="cpu", max_batch=max_batch)
    ref = _reference_loop(wav, m)
    assert list(got.keys()) == list(ref.keys())
    for k in ref:
        assert got[k].shape == ref[k].shape, k
        assert torch.equal(got[k], ref[k]), k


def test_session_shorter_than_one_window_raises():
    from voiceactivityprojection_b200.session import step_extraction

    with pytest.raises(RuntimeError, match="maximum size for tensor"):
        step_extraction(torch.zeros(1, 2, 16000 * 20), _FakeModel(), device="cpu")


def test_load_waveform_int16_scaling_and_resample(tmp_path):
    from scipy.io import wavfile
    import torchaudio.functional as AF

    from voiceactivityprojection_b200.audio import load_waveform

    rng = np.random.default_rng(0)
    pcm = (rng.standard_normal(24000) * 3000).astype(np.int16)
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 24000, pcm)
    x, sr = load_waveform(p, sample_rate=None)
    assert sr == 24000 and x.shape == (1, 24000)
    d, sr = load_waveform(p)  # the reference's default: resampled to 16 kHz (vap/audio.py:41)
    assert sr == 16000 and d.shape == (1, 16000)
    assert torch.equal(x[0], torch.from_numpy(pcm.astype(np.float32)) / 32768.0)
    y, sr = load_waveform(p, sample_rate=16000)
    assert sr == 16000 and y.shape == (1, 16000)
    assert torch.equal(y, AF.resample(x, 24000, 16000))
    st = np.stack([pcm, -pcm], axis=1)
    wavfile.write(p, 16000, st)
    z, _ = load_waveform(p, sample_rate=16000)
    assert z.shape == (2, 24000) and torch.equal(z[0], -z[1])
    zm, _ = load_waveform(p, None, None, None, True)  # positional order of the reference: ..., end_time, mono
    assert zm.shape == (1, 24000) and zm.abs().max() == 0


def test_config_argparse_mirror_and_cli_flags():
    from voiceactivityprojection_b200 import VapConfig
    from voiceactivityprojection_b200.run import get_args

    args, conf = get_args(["-a", "x.wav", "-sd", "s.pt", "--vap_channel_layers", "2", "--chunk"])
    assert conf == VapConfig(channel_layers=2) and args.chunk and args.filename is None
    assert conf.bin_times == [0.2, 0.4, 0.6, 0.8] and conf.dim == 256 and conf.cross_layers == 3


def test_json_contract(tmp_path):
    from voiceactivityprojection_b200.utils import read_json, tensor_dict_to_json, write_json

    out = {k: torch.arange(6, dtype=torch.float32).reshape(1, 3, 2) for k in ["probs", "vad", "p_now", "p_future"]}
    out["H"] = torch.zeros(1, 3)
    out["loss"] = torch.ones(1, 1)
    p = str(tmp_path / "o.json")
    write_json(tensor_dict_to_json(out), p)
    d = read_json(p)
    assert list(d.keys()) == ["probs", "vad", "p_now", "p_future", "H", "loss"]
    assert d["p_now"] == [[[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]]]
    assert json.load(open(p)) == d


def test_vad_postprocessing_run_length_filters():
    from voiceactivityprojection_b200.utils import vad_fill_silences, vad_omit_spikes

    v = torch.tensor([[1, 0], [0, 0], [1, 1], [1, 0], [0, 0], [0, 1], [1, 0]], dtype=torch.float32)
    f = vad_fill_silences(v.clone(), max_fill_time=0.02, frame_hz=50)  # 1-frame silences become active
    assert f[:, 0].tolist() == [1, 1, 1, 1, 0, 0, 1]
    o = vad_omit_spikes(v.clone(), max_omit_time=0.02, frame_hz=50)    # 1-frame activity removed
    assert o[:, 1].tolist() == [0, 0, 0, 0, 0, 0, 0]
    assert o[:, 0].tolist() == [0, 0, 1, 1, 0, 0, 0]


def test_load_state_dict_strict_false_drops_foreign_keys():
    from oracle import synth
    from voiceactivityprojection_b200 import VapConfig, VapGPT

    sd = dict(synth.make_state_dict(0))
    sd["VAP.codebook.emb.weight"] = torch.zeros(256, 8)
    sd["some_callback.state"] = torch.zeros(3)
    m = VapGPT(VapConfig())
    m.load_state_dict(sd, strict=False)
    assert "some_callback.state" not in m.state_dict() and "vap_head.weight" in m.state_dict()


def _reference_extractor_loop(wav, model, skip_last):
    """vap/extraction.py:182-258 restated: range(1, len(folds[1:])) skips the last unfold window."""
    cs, ss, sf = 400000, 80000, 250
    folds = wav.unfold(dimension=-1, size=cs, step=ss).permute(2, 0, 1, 3)
    out = model.probs(folds[0])
    keys = ["vad", "p_now", "p_future", "probs", "H"]
    stop = len(folds[1:]) if skip_last else len(folds)
    for ii in range(1, stop):
        o = model.probs(folds[ii])
        for k in keys:
            out[k] = torch.cat([out[k], o[k][:, -sf:]], dim=1)
    expected = round(round(wav.shape[-1] / 16000, 2) * 50)
    if expected != out["p_now"].shape[1]:
        om = expected - out["p_now"].shape[1]
        o = model.probs(wav[..., -cs:])
        for k in keys:
            out[k] = torch.cat([out[k], o[k][:, -om:]], dim=1)
    return out


@pytest.mark.parametrize("skip_last", [False, True])
def test_vap_extractor_stitching_and_minimal_outputs(tmp_path, skip_last):
    from voiceactivityprojection_b200.extraction import VapExtractor, get_minimal_output_json, write_minimal_csv

    m = _FakeModel()
    m.probs_orig = m.probs
    ex = VapExtractor(model=m, max_batch=3, compat_skip_last_fold=skip_last)
    ex.device = "cpu"
    wav = torch.randn(1, 2, int(47.3 * 16000), generator=torch.Generator().manual_seed(3))
    got = ex.step_extraction(wav)
    ref = _reference_extractor_loop(wav, m, skip_last)
    for k in ref:
        assert torch.equal(got[k], ref[k]), k
    mo = get_minimal_output_json(got, vad=None)
    assert list(mo.keys()) == ["p_now", "p_future", "model_vad0", "model_vad1", "H", "loss"]
    assert len(mo["p_now"]) == got["p_now"].shape[1] and len(mo["loss"]) == got["loss"].shape[1]
    p = str(tmp_path / "min.csv")
    write_minimal_csv(mo, p)
    lines = open(p).read().strip().splitlines()
    assert lines[0] == "p_now,p_future,model_vad0,model_vad1,H,loss" and len(lines) == 1 + len(mo["p_now"])
    assert lines[-1].endswith(",0")  # loss is shorter than the frame axis: padded with 0 like json_data_to_df


def test_zero_shot_mirror_builds_the_reference_subsets():
    """voiceactivityprojection_b200.zero_shot.ZeroShot: same class subsets (values and order) as the reference's
    (tests/golden/zero_shot.npz), packed into vapb_zero_shot's ten 256-bit sets; CPU tensors are refused."""
    from conftest import GOLDEN_DIR
    from voiceactivityprojection_b200.zero_shot import ZeroShot

    g = np.load(os.path.join(GOLDEN_DIR, "zero_shot.npz"))
    zs = ZeroShot(bin_times=[0.2, 0.4, 0.6, 0.8], frame_hz=50)
    names = ["subset_silence", "subset_silence_hold", "subset_active", "subset_active_hold", "bc_prediction"]
    for k in names:
        assert np.array_equal(getattr(zs, k).numpy(), g[k]), k
    for gi, k in enumerate(names):
        for s in (0, 1):
            words = [zs._sets[(2 * gi + s) * 8 + w] for w in range(8)]
            members = [c for c in range(256) if (words[c // 32] >> (c % 32)) & 1]
            assert members == sorted(g[k][s].tolist())
    with pytest.raises(RuntimeError):
        zs.get_probs(torch.zeros(1, 4, 256), torch.zeros(1, 4, 2))
    with pytest.raises(NotImplementedError):
        ZeroShot(bin_times=[0.2, 0.4], frame_hz=50)


def _run_bench(args, env_extra=None, timeout=300):
    import subprocess
    import sys

    env = dict(os.environ, **(env_extra or {}))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return subprocess.run([sys.executable, os.path.join(root, "bench.py"), *args], capture_output=True, text=True,
                          env=env, timeout=timeout, cwd=root)


def test_bench_reference_arm_prints_one_contract_line_and_only_on_rank_0():
    """`bench.py --impl reference` (the CPU oracle port timed on the host cores): exactly one JSON line on stdout
    with the contract's keys; under a multi-rank launch only rank 0 works and prints."""
    r = _run_bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-batch", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "1 chunks of 20 s" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r1 = _run_bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                    {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_bench_own_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a machine without CUDA the product arm exits non-zero instead of timing something else."""
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    r = _run_bench(["--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-e2e"])
    assert r.returncode != 0
    assert r.stdout.strip() == ""


def test_vad_list_helpers_match_reference_golden():
    """utils.get_vad_list_subset / vad_list_to_onehot / vad_onehot_to_vad_list / get_dialog_states / add_zero_channel
    against the unmodified reference's outputs (tests/golden/vad_list.json, oracle/make_golden_vadlist.py):
    identical JSON (values and int/float types) and identical tensors."""
    from conftest import GOLDEN_DIR
    from voiceactivityprojection_b200 import utils as U

    g = json.load(open(os.path.join(GOLDEN_DIR, "vad_list.json")))
    for c in g["subset"]:
        assert json.dumps(U.get_vad_list_subset(c["vad_list"], c["start"], c["end"])) == json.dumps(c["out"]), c
    for c in g["onehot_to_list"]:
        got = U.vad_onehot_to_vad_list(torch.tensor(c["vad"]).float(), c["frame_hz"], c["ipu_thresh_time"])
        assert json.dumps(got) == json.dumps(c["out"])
    for c in g["list_to_onehot"]:
        got = U.vad_list_to_onehot(c["vad_list"], c["duration"], **c["kw"])
        assert got.dtype == torch.float32 and torch.equal(got.long(), torch.tensor(c["out"]))
    d = g["dialog_states"]
    assert U.get_dialog_states(torch.tensor(d["vad"]).float()).tolist() == d["out"]
    w = torch.randn(2, 1, 10)
    z = U.add_zero_channel(w)
    assert z.shape == (2, 2, 10) and torch.equal(z[:, :1], w) and not z[:, 1].any()
    with pytest.raises(AssertionError):
        U.vad_list_to_onehot([[], []], 1.0)
    with pytest.raises(AssertionError):
        U.vad_onehot_to_vad_list(torch.zeros(5, 2))


def test_audio_header_helpers(tmp_path):
    """audio.get_audio_info / time_to_frames / sample_to_time (vap/audio.py:14-36) on a 24 kHz stereo int16 file."""
    import scipy.io.wavfile

    from voiceactivityprojection_b200 import audio as A

    path = str(tmp_path / "x.wav")
    scipy.io.wavfile.write(path, 24000, np.zeros((36000, 2), dtype=np.int16))
    info = A.get_audio_info(path)
    assert info == {"name": path, "duration": 1.5, "sample_rate": 24000, "num_frames": 36000, "bits_per_sample": 16,
                    "num_channels": 2, "encoding": "PCM_S"}
    assert A.time_to_frames(0.999, 0.02) == 49 and A.time_to_samples(0.5, 16000) == 8000
    assert A.sample_to_time(8000, 16000) == 0.5


def test_extraction_cli_writes_minimal_json_and_csv(tmp_path):
    """`python -m voiceactivityprojection_b200.extraction` (vap/extraction.py:18-56, 340-378): flags, vad-list input,
    `<audio name>.json` / `.csv` with the minimal keys, on a model double (the real model needs the GPU)."""
    import scipy.io.wavfile

    from voiceactivityprojection_b200 import extraction as E

    wav_path = str(tmp_path / "dialog.wav")
    pcm = (np.random.default_rng(0).standard_normal((16000 * 12, 2)) * 3000).astype(np.int16)
    scipy.io.wavfile.write(wav_path, 16000, pcm)
    vad_path = str(tmp_path / "dialog_vad_list.json")
    json.dump([[[0.5, 2.0], [6.0, 7.5]], [[2.5, 5.0]]], open(vad_path, "w"))
    m = _FakeModel()
    m.device = "cpu"
    out = E.main(["-a", wav_path, "-v", vad_path, "--output_dir", str(tmp_path)], model=m)
    assert out == str(tmp_path / "dialog.json")
    d = json.load(open(out))
    assert list(d.keys()) == ["p_now", "p_future", "model_vad0", "model_vad1", "H", "loss", "vad0", "vad1"]
    assert len(d["p_now"]) == 600 and len(d["loss"]) == 500 and len(d["vad0"]) == 600
    assert d["vad0"][24] == 0.0 and d["vad0"][25] == 1.0 and d["vad1"][249] == 1.0 and d["vad1"][250] == 0.0
    out = E.main(["-a", wav_path, "--output_format", "csv", "--output_dir", str(tmp_path)], model=m)
    lines = open(out).read().strip().splitlines()
    assert out.endswith("dialog.csv") and lines[0] == "p_now,p_future,model_vad0,model_vad1,H,loss" and len(lines) == 601
    df = E.json_data_to_df({k: v for k, v in d.items() if not k.startswith("vad")})
    assert df.shape == (600, 6) and df["loss"].iloc[-1] == 0
    args, conf = E.get_args(["-a", "x.wav", "--context_time", "10", "--step_time", "2.5"])
    assert args.context_time == 10 and args.step_time == 2.5 and args.output_format == "json" and conf.frame_hz == 50


def test_objective_helper_classes_match_reference_golden():
    """objective.ProjectionWindow / Codebook / get_da_labels / loss_vad against the unmodified reference's outputs
    (tests/golden/objective.npz, oracle/make_golden_objective.py): integer results bit-exact."""
    from conftest import GOLDEN_DIR
    from voiceactivityprojection_b200.objective import ObjectiveVAP

    g = np.load(os.path.join(GOLDEN_DIR, "objective.npz"))
    o = ObjectiveVAP()
    va = torch.from_numpy(g["va"]).float()
    idx, ds = o.get_da_labels(va)
    assert torch.equal(idx, torch.from_numpy(g["labels"])) and torch.equal(ds, torch.from_numpy(g["dialog_states"]))
    assert torch.equal(o.get_labels(va), idx)
    wins = o.projection_window_extractor(va)
    assert wins.dtype == torch.float32 and torch.equal(wins, torch.from_numpy(g["windows"]).float())
    assert o.projection_window_extractor.projection(va).shape == (3, 160, 2, 100)
    assert repr(o.projection_window_extractor) == str(g["repr_pw"])
    cb = o.codebook
    assert (cb.n_bins, cb.total_bins, cb.n_classes) == (4, 8, 256)
    assert torch.equal(cb.emb.weight, torch.from_numpy(g["code_vectors"]))
    assert torch.equal(cb.encode(torch.from_numpy(g["soft"])), torch.from_numpy(g["soft_idx"]))
    assert torch.equal(cb.decode(torch.from_numpy(g["some_idx"])), torch.from_numpy(g["some_windows"]))
    assert torch.equal(cb(cb.decode(torch.arange(256))), torch.arange(256))
    assert torch.equal(cb.single_idx_to_onehot(37), cb.emb.weight[37])
    lv = o.loss_vad(torch.from_numpy(g["vad_logits"]), torch.from_numpy(g["vad_target"]))
    assert abs(float(lv) - float(g["loss_vad"])) <= 1e-6
    with pytest.raises(AssertionError):
        cb.encode(torch.zeros(3, 2, 5))
