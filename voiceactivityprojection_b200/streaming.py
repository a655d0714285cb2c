"""Rolling-window streaming inference: the model side of the reference's realtime loop
`sds/run_sds.py` (`TurnTakingSDS.add_audio_bytes_to_tensor` :206-220 and the body of `run` :232-247),
without the audio capture (PyAudio) and the ZeroMQ publisher, which are outside the inference path.

The reference keeps a (1, 2, context*sample_rate) float tensor, rolls it left by every captured chunk
(interleaved stereo int16 bytes scaled by 1/32768), recomputes the WHOLE window with `model.probs` at
every poll and publishes the mean of `p_now[:, -tt_frames:, 0]`. Here the ring lives on the device,
the int16 bytes cross PCIe as they are (half the bytes of float) and are de-interleaved and scaled
there; the full-window recompute is kept (same numerics as the reference), it takes ~4 ms on a B200
in the tensor modes.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

NORM_FACTOR = 1.0 / (2 ** 15)  # sds/run_sds.py: int16 -> [-1, 1)


class StreamingVAP:
    def __init__(self, model, context_time: float = 20.0, tt_time: float = 0.5, precision: Optional[str] = None):
        if model._device.type != "cuda":
            raise RuntimeError("StreamingVAP needs the model on a CUDA device (no CPU fallback)")
        self.model, self.precision = model, precision
        self.device = model._device
        self.n_samples = round(context_time * model.sample_rate)  # sds/run_sds.py:173
        self.tt_frames = round(tt_time * model.frame_hz)          # :182
        self.x = torch.zeros((1, 2, self.n_samples), dtype=torch.float32, device=self.device)

    def add_audio_bytes(self, audio_bytes: bytes) -> int:
        """Appends interleaved stereo int16 PCM (what the PyAudio callback of the reference delivers)."""
        pcm = torch.frombuffer(bytearray(audio_bytes), dtype=torch.int16)
        return self.add_audio_int16(pcm)

    def add_audio_int16(self, pcm: Tensor) -> int:
        """pcm: 1-D int16, samples interleaved (a0, b0, a1, b1, ...). Returns the frames appended per channel."""
        if pcm.numel() % 2:
            raise ValueError("interleaved stereo needs an even number of samples")
        n = pcm.numel() // 2
        if n == 0:
            return 0
        if n > self.n_samples:  # more than a window: only the newest window's worth matters
            pcm, n = pcm[-2 * self.n_samples:], self.n_samples
        d = pcm.to(self.device, non_blocking=True).view(n, 2).t().to(torch.float32) * NORM_FACTOR
        self.x = self.x.roll(-n, -1)
        self.x[0, :, -n:] = d
        return n

    @torch.no_grad()
    def step(self, levels: bool = False) -> Dict[str, object]:
        """One poll of the reference's loop: full-window probs and the scalar it publishes (:241-242).
        levels=True adds the two integers the loop prints next to it (:236-237): 100 x the peak |sample| of each
        speaker over the newest 4000 samples."""
        kw = {} if self.precision is None else {"precision": self.precision}
        out = self.model.probs(self.x, **kw)
        ret = {"p_now_mean": out["p_now"][0, -self.tt_frames:, 0].mean().item(), "out": out}
        if levels:
            peak = (self.x[0, :, -4000:].abs().amax(-1) * 100).long().tolist()
            ret["level_a"], ret["level_b"] = peak
        return ret
