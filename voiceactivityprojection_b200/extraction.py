"""File-level extraction: the reference's `vap/extraction.py` (`VapExtractor` :99-270,
`get_minimal_output_json` :82-96, `json_data_to_df` :63-79) on top of the batched window scheduler
of `session.py`.

Two defects of the reference are fixed here and kept reproducible behind a flag:
  * `VapExtractor.step_extraction` iterates `range(1, len(folds[1:]))` (:211) and so SKIPS THE LAST
    unfold window; the frames it would have contributed are then (silently) taken from the
    right-aligned tail window instead. `compat_skip_last_fold=True` reproduces that stitching.
  * `VapExtractor.extract` calls `self.model(waveform, vad=vad)` for short files (:268), which
    `VapGPT.forward` does not accept (TypeError in the reference). Short files go through `probs` here.
Ground-truth `vad` passed by the caller is only echoed into the minimal JSON (`vad0` / `vad1`): the
reference's `probs` overwrites it with the model's own VAD before computing anything.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import Tensor

from .session import STITCH_KEYS, window_plan

STEP_EXTRACTION_LIMIT = 160  # seconds (vap/extraction.py: longer files are processed in 25 s windows)


def get_minimal_output_json(out: Dict[str, Tensor], vad: Optional[Tensor] = None) -> Dict[str, list]:
    min_out = {
        "p_now": out["p_now"][0, :, 0].tolist(),
        "p_future": out["p_future"][0, :, 0].tolist(),
        "model_vad0": out["vad"][0, :, 0].tolist(),
        "model_vad1": out["vad"][0, :, 1].tolist(),
        "H": out["H"][0].tolist(),
    }
    if "loss" in out:
        min_out["loss"] = out["loss"][0].tolist()
    if vad is not None:
        min_out["vad0"] = vad[0, :, 0].tolist()
        min_out["vad1"] = vad[0, :, 1].tolist()
    return min_out


def minimal_output_rows(min_out: Dict[str, list]) -> List[Dict[str, float]]:
    """One row per frame (what `json_data_to_df` feeds to pandas): shorter columns (`loss`) are padded with 0."""
    n = len(min_out["p_now"])
    return [{k: (v[i] if i < len(v) else 0) for k, v in min_out.items()} for i in range(n)]


def write_minimal_csv(min_out: Dict[str, list], path: str) -> None:
    import csv

    rows = minimal_output_rows(min_out)
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(min_out.keys()))
        w.writeheader()
        w.writerows(rows)


class VapExtractor:
    """
    input: |------------ chunk time ---------------------|
    input: |------ context time -------|--- step time ---|
    """

    def __init__(self, context_time: float = 20, step_time: float = 5, state_dict_path: Optional[str] = None,
                 model=None, precision: Optional[str] = None, max_batch: int = 64, compat_skip_last_fold: bool = False):
        if model is None:
            from .model import VapConfig, VapGPT

            model = VapGPT(VapConfig(), precision=precision)
            model.load_state_dict(torch.load(state_dict_path, map_location="cpu"))
            model = model.to("cuda").eval()
        self.model = model
        self.device = getattr(model, "device", "cuda")  # VapGPT lives on the GPU; test doubles may say otherwise
        self.precision, self.max_batch, self.compat_skip_last_fold = precision, max_batch, compat_skip_last_fold
        self.context_time, self.step_time = context_time, step_time
        self.chunk_time = context_time + step_time
        self.step_samples = int(step_time * model.sample_rate)
        self.chunk_samples = int(self.chunk_time * model.sample_rate)
        self.step_frames = int(step_time * model.frame_hz)
        self.chunk_frames = int(self.chunk_time * model.frame_hz)

    def __repr__(self):
        return (f"VapExtractor\nContext time: {self.context_time}s\nStep time: {self.step_time}s\n"
                f"Chunk time: {self.chunk_time}s\nStep samples: {self.step_samples}\nChunk samples: {self.chunk_samples}\n"
                f"Step frames: {self.step_frames}\nChunk frames: {self.chunk_frames}\n")

    def _probs(self, w: Tensor) -> Dict[str, Tensor]:
        kw = {} if self.precision is None else {"precision": self.precision}
        return self.model.probs(w, **kw)

    @torch.no_grad()
    def step_extraction(self, waveform: Tensor, vad: Optional[Tensor] = None, **_ignored) -> Dict[str, Tensor]:
        plan = window_plan(waveform.shape[-1], self.model.sample_rate, self.model.frame_hz, self.context_time, self.step_time)
        B, cs, sf = waveform.shape[0], self.chunk_samples, self.step_frames
        wav = waveform.to(self.device)
        folds = wav.unfold(dimension=-1, size=cs, step=self.step_samples).permute(2, 0, 1, 3)
        nf = folds.shape[0] - (1 if self.compat_skip_last_fold and folds.shape[0] > 1 else 0)
        out = {k: v.clone() for k, v in self._probs(folds[0].contiguous()).items()}
        tails = {k: [] for k in STITCH_KEYS}
        per = max(1, self.max_batch // B)
        for i in range(1, nf, per):
            j = min(nf, i + per)
            o = self._probs(folds[i:j].reshape((j - i) * B, 2, cs))
            for k in STITCH_KEYS:
                v = o[k][:, -sf:]
                v = v.reshape(j - i, B, *v.shape[1:]).transpose(0, 1)
                tails[k].append(v.reshape(B, (j - i) * sf, *v.shape[3:]))
        for k in STITCH_KEYS:
            out[k] = torch.cat([out[k]] + tails[k], dim=1)
        processed = out["p_now"].shape[1]
        if plan["expected_frames"] != processed:
            omitted = plan["expected_frames"] - processed
            assert omitted < self.chunk_frames, f"Omitted frames {omitted} > chunk frames {self.chunk_frames}"
            o = self._probs(wav[..., -cs:].contiguous())
            for k in STITCH_KEYS:
                out[k] = torch.cat([out[k], o[k][:, -omitted:]], dim=1)
        return {k: v.cpu() for k, v in out.items()}

    @torch.no_grad()
    def extract(self, waveform: Tensor, vad: Optional[Tensor] = None) -> Dict[str, Tensor]:
        duration = waveform.shape[-1] / self.model.sample_rate
        if duration > STEP_EXTRACTION_LIMIT:
            return self.step_extraction(waveform, vad=vad)
        return {k: v.cpu() for k, v in self._probs(waveform.to(self.device)).items()}


# --------------------------------------------------------------------------- #
# command line (vap/extraction.py:18-56, 340-378)                              #
# --------------------------------------------------------------------------- #
def get_duration(x: Tensor, sample_rate: int = 16000) -> float:
    return x.shape[-1] / sample_rate


def json_data_to_df(out: Dict[str, list]):
    """vap/extraction.py:63-79: one DataFrame row per frame, short columns padded with 0."""
    import pandas as pd

    return pd.DataFrame(minimal_output_rows(out))


def get_args(argv=None):
    from argparse import ArgumentParser

    from .model import VapConfig

    parser = ArgumentParser()
    parser.add_argument("-a", "--audio", type=str, help="Path to waveform")
    parser.add_argument("-v", "--vad", type=str, help="Path to vad list", default=None)
    parser.add_argument("--output_format", type=str, default="json", help="output format: ['json', 'csv']. Default='json'")
    parser.add_argument("-sd", "--state_dict", type=str,
                        default="example/VAP_3mmz3t0u_50Hz_ad20s_134-epoch9-val_2.56.pt", help="Path to state_dict")
    parser, _ = VapConfig.add_argparse_args(parser, [])
    parser.add_argument("--context_time", type=float, default=20, help="Duration of each chunk processed by model")
    parser.add_argument("--step_time", type=float, default=5, help="Increment to process in a step")
    parser.add_argument("--precision", default=None, choices=["fp32", "bf16", "fp16"])
    parser.add_argument("--output_dir", type=str, default=".", help="where <audio name>.json / .csv is written")
    args = parser.parse_args(argv)
    return args, VapConfig.args_to_conf(args)


def main(argv=None, model=None) -> str:
    """`python -m voiceactivityprojection_b200.extraction -a X.wav -sd S.pt [-v X_vad_list.json]`: the reference's
    `python vap/extraction.py` — minimal JSON (or CSV) next to the audio's base name. Returns the written path.
    `model`: an already loaded VapGPT (tests / callers that hold one)."""
    import os

    from .audio import load_waveform
    from .utils import read_json, vad_list_to_onehot, write_json

    args, _ = get_args(argv)
    extractor = VapExtractor(context_time=args.context_time, step_time=args.step_time,
                             state_dict_path=args.state_dict, model=model, precision=args.precision)
    sr = extractor.model.sample_rate
    waveform, _ = load_waveform(args.audio, sample_rate=sr)
    if waveform.shape[0] == 1:  # run.py:219-220: a silent second speaker under a mono file
        waveform = torch.cat((waveform, torch.zeros_like(waveform)))
    waveform = waveform.unsqueeze(0)
    print("waveform: ", tuple(waveform.shape))
    vad = None
    if args.vad is not None:
        vad = vad_list_to_onehot(read_json(args.vad), duration=get_duration(waveform, sr),
                                 frame_hz=extractor.model.frame_hz).unsqueeze(0)
        print("vad: ", tuple(vad.shape))
    min_out = get_minimal_output_json(extractor.extract(waveform), vad)
    print("Keys:    Frames")
    for k, v in min_out.items():
        print(f"{k}:     {len(v)}")
    base = os.path.join(args.output_dir, os.path.basename(args.audio).replace(".wav", ""))
    if args.output_format == "json":
        path = base + ".json"
        write_json(min_out, path)
    else:
        path = base + ".csv"
        write_minimal_csv(min_out, path)
    print("Saved -> ", path)
    return path


if __name__ == "__main__":
    main()
