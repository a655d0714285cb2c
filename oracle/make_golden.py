"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the authoring container only (needs /root/reference):

    python -m oracle.make_golden

Each fixture holds the inputs' recipe (seed / kind / shape, or the resampled
example wav itself) and the outputs of the reference's own `VapGPT.forward` /
`VapGPT.probs` (vap/model.py:180-268) and `run.py:step_extraction` on CPU fp32,
with weights from oracle.synth.make_state_dict (the shipped checkpoints are
missing). The GPU box re-creates weights and inputs from the
recipe and compares the CUDA path with these arrays.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch

from oracle import ref_import, synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> recipe
CASES = {
    "lstm1_turns_T125": dict(seed=1, ar_mode="LSTM", ar_layers=1, gain=2.0, batch=2,
                             n_samples=40000, wav_seed=0, kind="turns", stages=True),
    "gru1_noise_T117": dict(seed=2, ar_mode="GRU", ar_layers=1, gain=2.0, batch=1,
                            n_samples=37392, wav_seed=1, kind="noise"),
    "lstm2_turns_T103": dict(seed=3, ar_mode="LSTM", ar_layers=2, gain=2.0, batch=1,
                             n_samples=33000, wav_seed=2, kind="turns"),
    "lstm1_mono_T500": dict(seed=4, ar_mode="LSTM", ar_layers=1, gain=1.0, batch=1,
                            n_samples=160000, wav_seed=3, kind="mono", no_probs=True),
    "lstm1_turns_T1000": dict(seed=5, ar_mode="LSTM", ar_layers=1, gain=2.0, batch=1,
                              n_samples=320000, wav_seed=4, kind="turns", no_probs=True),
}


def load_example_wav():
    """What run.py:217-221 feeds the model for the example file, without
    torchaudio.load (needs TorchCodec here): int16/32768 then AF.resample
    (vap/audio.py:65-68), zero second channel, batch dim."""
    import scipy.io.wavfile
    import torchaudio.functional as AF

    path = os.path.join(ref_import.REF_ROOT, "example", "student_long_female_en-US-Wavenet-G.wav")
    sr, d = scipy.io.wavfile.read(path)
    x = torch.from_numpy(d.astype(np.float32) / 32768.0)[None]
    x = AF.resample(x, orig_freq=sr, new_freq=16000)
    x = torch.cat((x, torch.zeros_like(x)))
    return x.unsqueeze(0)


def reference_step_extraction():
    """run.py cannot be imported (matplotlib). Take the `step_extraction`
    FunctionDef out of its source, unmodified, and compile it alone."""
    src = open(os.path.join(ref_import.REF_ROOT, "run.py")).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "step_extraction"][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    sys.path.insert(0, ref_import.REF_ROOT)
    from vap.utils import batch_to_device

    ns = {"torch": torch, "batch_to_device": batch_to_device}
    exec(compile(mod, "run.py:step_extraction", "exec"), ns)
    return ns["step_extraction"]


def run_case(name, c):
    sd = synth.make_state_dict(c["seed"], c["ar_mode"], c["ar_layers"], c["gain"])
    model = ref_import.build_reference(sd, c["ar_mode"], c["ar_layers"])
    if c.get("kind") == "example":
        wav = load_example_wav()
    else:
        wav = synth.make_waveform(c["batch"], c["n_samples"], c["wav_seed"], c["kind"])
    out = {}
    stages = {}
    hooks = []
    if c.get("stages"):
        def grab(key):
            def f(mod, inp, o):
                if key not in stages:  # first call = channel 0
                    stages[key] = (o["x"] if isinstance(o, dict) else o).detach().clone()
            return f

        enc = model.encoder
        hooks.append(enc.encoder.gEncoder.register_forward_hook(grab("conv_ncw")))
        hooks.append(enc.encoder.gAR.register_forward_hook(grab("ar")))
        hooks.append(enc.downsample.register_forward_hook(grab("enc")))
        hooks.append(model.ar_channel.register_forward_hook(grab("ch")))
        for l, layer in enumerate(model.ar.layers):
            def f(mod, inp, o, l=l):
                stages[f"ar{l}_x1"], stages[f"ar{l}_x2"] = o[0].detach().clone(), o[1].detach().clone()
            hooks.append(layer.register_forward_hook(f))
        hooks.append(model.ar.combinator.register_forward_hook(grab("comb")))
    with torch.no_grad():
        fwd = model(wav)
    for h in hooks:
        h.remove()
    pr = model.probs(wav)
    out["logits"] = fwd["logits"]
    out["vad_logits"] = fwd["vad"]
    for k, v in pr.items():
        if k == "probs" and c.get("no_probs"):
            continue
        out[k] = v
    for k, v in stages.items():
        if k == "conv_ncw":
            out["stage_conv"] = v[:1].transpose(1, 2).contiguous()  # (1, T100, 256), channel 0
        elif k.startswith("ar") and k[2:3].isdigit() or k == "comb":
            out["stage_" + k] = v[:1]
        else:
            out["stage_" + k] = v[:1]
    arrays = {k: v.numpy() for k, v in out.items()}
    if c.get("kind") == "example":
        arrays["waveform"] = wav.numpy()
    arrays["recipe"] = np.array(repr(c))
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB", {k: v.shape for k, v in arrays.items() if k != 'recipe'})


def run_session(name="session_45s"):
    """run.py:23-131 on a 45 s session (5 unfold windows of 25 s + no tail) and
    a 47.3 s one (tail window path)."""
    step_extraction = reference_step_extraction()
    c = dict(seed=6, ar_mode="LSTM", ar_layers=1, gain=2.0)
    sd = synth.make_state_dict(c["seed"], c["ar_mode"], c["ar_layers"], c["gain"])
    model = ref_import.build_reference(sd)
    arrays = {}
    for tag, n in [("a", 720000), ("b", 756800)]:
        wav = synth.make_waveform(1, n, 7, "turns")
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            out = step_extraction(wav, model, "cpu", pbar=False)
        for k in ["vad", "p_now", "p_future", "H", "loss"]:
            arrays[f"{tag}_{k}"] = out[k].numpy()
        arrays[f"{tag}_probs_argmax"] = out["probs"].argmax(-1).to(torch.uint8).numpy()
        arrays[f"{tag}_probs_max"] = out["probs"].max(-1).values.numpy()
        arrays[f"{tag}_n_samples"] = np.array(n)
    c.update(wav_seed=7, kind="turns")
    arrays["recipe"] = np.array(repr(c))
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB", {k: v.shape for k, v in arrays.items() if k != 'recipe'})


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    for name, c in CASES.items():
        run_case(name, c)
    run_case("example_wav_T117", dict(seed=0, ar_mode="LSTM", ar_layers=1, gain=2.0, kind="example"))
    run_session()


if __name__ == "__main__":
    main()
