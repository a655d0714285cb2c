"""Drop-in facade for the reference's `vap.model` on the stereo inference path.

Mirrors `VapConfig` and `VapGPT` of the reference (vap/model.py:42-79, 125-268)
— same constructor, attributes (`sample_rate`, `frame_hz`, `conf`, `objective`,
`horizon_time`), `load_state_dict` of the reference's flat state dicts,
`forward`, `probs`, `vad`, `.to()/.eval()` — but every device op is a hand-written
sm_100a kernel behind the C-ABI of include/vapb.h (libvapb.so, loaded with
ctypes). There is no CPU implementation: CPU tensors raise.

Differences a caller can see:
  * `VapGPT()` needs no CPC checkpoint; the gAR cell (LSTM/GRU) and depth are
    inferred from the state dict in `load_state_dict`.
  * `forward(..., attention=True)` returns the reference's attention maps from a
    separate fp32 map kernel (the fused attention kernels never materialise them).
  * `precision="fp32"` (default; CUDA-core FMA, parity with the reference),
    `"fp32_tc"` (the same fp32 path with every GEMM on the tensor cores at
    fp32-class accuracy, 2.7x faster), `"bf16"` or `"fp16"` (tcgen05 tensor cores;
    same kernels and speed, fp16 operands give ~8x lower error) — constructor
    keyword, attribute, or env VAPB_PRECISION.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from .objective import ObjectiveVAP
from .utils import vad_fill_silences, vad_omit_spikes

BIN_TIMES: list = [0.2, 0.4, 0.6, 0.8]


def load_older_state_dict(path="example/VAP_3mmz3t0u_50Hz_ad20s_134-epoch9-val_2.56.ckpt"):
    """Lightning .ckpt -> flat state dict (reference vap/model.py:28-39)."""
    sd = torch.load(path, map_location="cpu", weights_only=False)["state_dict"]
    new_sd = {}
    for k, v in sd.items():
        if "VAP.codebook" in k:
            continue
        if "vap_head" in k:
            k = k.replace("vap_head.projection_head", "vap_head")
        new_sd[k.replace("net.", "")] = v
    return new_sd


@dataclass
class VapConfig:
    """Field-for-field the reference's VapConfig (vap/model.py:42-79)."""

    sample_rate: int = 16_000
    frame_hz: int = 50
    bin_times: List[float] = field(default_factory=lambda: BIN_TIMES)

    # Encoder (training flags in the reference; accepted and ignored here)
    freeze_encoder: int = 1
    load_pretrained: int = 1

    # GPT
    dim: int = 256
    channel_layers: int = 1
    cross_layers: int = 3
    num_heads: int = 4
    dropout: float = 0.1

    @staticmethod
    def add_argparse_args(parser, fields_added=[]):
        for k, v in VapConfig.__dataclass_fields__.items():
            if k == "bin_times":
                parser.add_argument(f"--vap_{k}", nargs="+", type=float, default=v.default_factory())
            else:
                parser.add_argument(f"--vap_{k}", type=v.type if callable(v.type) else eval(v.type),
                                    default=v.default)
            fields_added.append(k)
        return parser, fields_added

    @staticmethod
    def args_to_conf(args):
        return VapConfig(
            **{k.replace("vap_", ""): v for k, v in vars(args).items() if k.startswith("vap_")}
        )


class VapGPT(nn.Module):
    def __init__(self, conf: Optional[VapConfig] = None, precision: Optional[str] = None):
        super().__init__()
        if conf is None:
            conf = VapConfig()
        self.conf = conf
        self.sample_rate = conf.sample_rate
        self.frame_hz = conf.frame_hz
        if conf.dim != 256 or conf.num_heads != 4 or conf.sample_rate != 16000 or conf.frame_hz != 50:
            raise NotImplementedError(
                "the B200 path implements the shipped configuration: dim=256, num_heads=4, 16 kHz, 50 Hz"
            )
        if [round(b * conf.frame_hz) for b in conf.bin_times] != [10, 20, 30, 40]:
            raise NotImplementedError("bin_times must be [.2, .4, .6, .8]")
        self.precision = precision or os.environ.get("VAPB_PRECISION", "fp32")
        if self.precision not in _lib.MODES:
            raise ValueError(f"precision must be one of {list(_lib.MODES)}")
        self.objective = ObjectiveVAP(bin_times=conf.bin_times, frame_hz=conf.frame_hz)
        self._sd: Dict[str, Tensor] = {}  # the reference-schema state dict (CPU fp32 master copy)
        self._device = torch.device("cpu")
        self._handle = None
        self._handle_device = None
        self._ws = {}
        self.objective._owner = self

    # ------------------------------------------------------------------ state
    @property
    def horizon_time(self):
        return self.objective.horizon_time

    def state_dict(self, *args, **kwargs):
        return {k: v.clone() for k, v in self._sd.items()}

    def load_state_dict(self, state_dict, strict: bool = True):
        """Strict load of a reference state dict (run.py:200-201). Layout (LSTM vs
        GRU, depth) is taken from the tensor shapes; errors are raised with the
        library's message, in nn.Module.load_state_dict's wording."""
        sd = {k: v.detach().to("cpu", torch.float32).contiguous() for k, v in state_dict.items()}
        if not strict:
            # nn.Module.load_state_dict(strict=False) ignores keys the module does not own (the reference loads older
            # Lightning checkpoints this way, vap/model.py:455); missing weights still cannot be invented
            known = ("encoder.", "ar_channel.", "ar.", "objective.codebook.", "va_classifier.", "vap_head.")
            sd = {k: v for k, v in sd.items() if k.startswith(known)}
        self._release()
        self._sd = sd
        if self._device.type == "cuda":
            self._ensure_handle()
        else:
            self._validate_on_host()
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def _validate_on_host(self):
        # Shape validation lives in the library (vapb_finalize), which needs a
        # device. Without one only the cheap structural facts are checked here.
        p = "encoder.encoder.gAR.baseNet.weight_ih_l0"
        if p not in self._sd:
            raise RuntimeError(f"Error(s) in loading state_dict for VapGPT: Missing key(s): {p}.")

    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0))
        if probe.dtype not in (torch.float32,):
            raise NotImplementedError("weights stay fp32; choose arithmetic with precision='bf16'")
        if probe.device != self._device:
            self._release()
            self._device = probe.device
        return self

    def eval(self):
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("inference-only implementation (VapGPT.forward/probs/vad)")
        return self

    def _release(self):
        if self._handle is not None:
            _lib.load().vapb_destroy(self._handle)
        self._handle = None
        self._ws = {}

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_handle(self):
        if self._handle is not None:
            return self._handle
        if self._device.type != "cuda":
            raise RuntimeError(
                "VapGPT (B200) runs on CUDA only; call model.to('cuda') — there is no CPU fallback"
            )
        if not self._sd:
            raise RuntimeError("no weights loaded: call load_state_dict() first")
        lib = _lib.load()
        dev = self._device.index if self._device.index is not None else torch.cuda.current_device()
        h = C.c_void_p()
        _lib.check(lib, None, lib.vapb_create(dev, C.byref(h)))
        try:
            for k, v in self._sd.items():
                shape = (C.c_int64 * max(v.ndim, 1))(*v.shape)
                _lib.check(lib, h, lib.vapb_load_tensor(h, k.encode(), v.data_ptr(), v.ndim, shape))
            _lib.check(lib, h, lib.vapb_finalize(h))
        except Exception:
            lib.vapb_destroy(h)
            raise
        self._handle, self._handle_device = h, dev
        # nn.Module.load_state_dict(strict=True) would refuse a checkpoint whose depth differs from the module built
        # from `conf`; the library infers the depth from the keys, so the comparison is made here (the
        # attention-map buffers of forward(attention=True) are sized from conf)
        d = self.describe()
        if (d["channel_layers"], d["cross_layers"], d["num_heads"]) != (
                self.conf.channel_layers, self.conf.cross_layers, self.conf.num_heads):
            self._release()
            raise RuntimeError(
                "Error(s) in loading state_dict for VapGPT: the state dict holds "
                f"{d['channel_layers']} ar_channel layer(s), {d['cross_layers']} ar layer(s) and {d['num_heads']} heads "
                f"but conf asks for channel_layers={self.conf.channel_layers}, cross_layers={self.conf.cross_layers}, "
                f"num_heads={self.conf.num_heads}")
        return h

    def describe(self) -> Dict[str, int]:
        lib, h = _lib.load(), self._ensure_handle()
        v = [C.c_int() for _ in range(5)]
        _lib.check(lib, h, lib.vapb_describe(h, *[C.byref(x) for x in v]))
        return dict(zip(["ar_kind", "ar_layers", "channel_layers", "cross_layers", "num_heads"],
                        [x.value for x in v]))

    def launch_count(self) -> int:
        lib, h = _lib.load(), self._ensure_handle()
        n = C.c_uint64()
        _lib.check(lib, h, lib.vapb_launch_count(h, C.byref(n)))
        return n.value

    # ------------------------------------------------------------------ helpers
    @property
    def device(self) -> torch.device:
        """The device the weights live on (the reference's callers read `model.device` / next(parameters()).device)."""
        return self._device

    def _check_input(self, waveform: Tensor, allow_pcm16: bool = False) -> Tensor:
        assert waveform.ndim == 3 and waveform.shape[1] == 2, (
            f"audio VAP ENCODER: {tuple(waveform.shape)} != (B, 2, n_samples)"
        )
        if waveform.device.type != "cuda":
            raise RuntimeError("waveform must be a CUDA tensor (no CPU fallback); use probs_host() for host buffers")
        if self._device.type != "cuda":
            self._device = waveform.device
        dev = self._device if self._device.index is not None else torch.device("cuda", torch.cuda.current_device())
        wdev = waveform.device if waveform.device.index is not None else torch.device("cuda", torch.cuda.current_device())
        if wdev != dev:
            # nn.Module semantics: inputs and parameters must share a device (the library's workspace, weights and
            # streams all belong to the model's device)
            raise RuntimeError(f"Expected all tensors to be on the same device, but found {wdev} (waveform) and "
                               f"{dev} (model)")
        if allow_pcm16 and waveform.dtype == torch.int16:
            return waveform.contiguous()
        return waveform.to(torch.float32).contiguous()

    def _pcm16_to_f32(self, pcm: Tensor) -> Tensor:
        """int16 PCM -> float32 / 32768 on the device (vapb_pcm16_to_f32), for the paths that cannot read PCM directly."""
        lib = _lib.load()
        out = torch.empty(pcm.shape, dtype=torch.float32, device=pcm.device)
        st = torch.cuda.current_stream(pcm.device).cuda_stream
        _lib.check(lib, None, lib.vapb_pcm16_to_f32(st, pcm.data_ptr(), pcm.numel(), out.data_ptr()))
        return out

    def _workspace(self, batch: int, n_samples: int, mode: int):
        lib, h = _lib.load(), self._ensure_handle()
        need = C.c_size_t()
        _lib.check(lib, h, lib.vapb_workspace_bytes(h, batch, n_samples, mode, C.byref(need)))
        # one scratch arena per (device, stream): calls on different streams may overlap, so they must not share it
        key = (str(self._device), torch.cuda.current_stream(self._device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need.value:
            self._ws.pop(key, None)
            ws = None  # release the old arena before allocating its replacement
            ws = self._ws[key] = torch.empty(need.value, dtype=torch.uint8, device=self._device)
        return ws

    def _mode(self, precision=None) -> int:
        return _lib.MODES[precision or self.precision]

    # ------------------------------------------------------------------ hot path
    @torch.no_grad()
    def encode_audio(self, audio: Tensor, precision: Optional[str] = None):
        """vap/model.py:169-175 -> (x1, x2), each (B, T, 256): the encoder output (CPC conv stack, gAR, 100 -> 50 Hz
        downsample) of the two speaker channels. forward() does not call this (the encoder is the first part of
        one fused call); it runs that call and reads the encoder stage back."""
        x = self.stage("enc", audio, precision=precision)
        B = audio.shape[0]
        return x[:B], x[B:]

    def vad_loss(self, vad_output: Tensor, vad: Tensor) -> Tensor:
        """vap/model.py:177-178: mean binary cross-entropy of the VAD logits (a metric on outputs; no device kernel)."""
        return torch.nn.functional.binary_cross_entropy_with_logits(vad_output, vad)

    @torch.no_grad()
    def forward(self, waveform: Tensor, attention: bool = False, precision: Optional[str] = None) -> Dict[str, Tensor]:
        """vap/model.py:249-268 -> {"logits": (B,T,256), "vad": (B,T,2)} (vad = logits).
        attention=True adds the reference's "self_attn" (B,2,channel_layers,H,T,T), "cross_attn" and
        "cross_self_attn" (B,2,cross_layers,H,T,T): a diagnostic output, computed by the fp32 path with a separate
        map kernel per attention (vapb_forward_attention); the fused attention kernels never materialise them."""
        wav = self._check_input(waveform)
        B, _, S = wav.shape
        lib, h = _lib.load(), self._ensure_handle()
        _, T = _lib.frames(S)
        if attention:
            if precision not in (None, "fp32"):
                raise ValueError("attention=True runs in fp32 mode")
            ws = self._workspace(B, S, _lib.MODE_FP32)
            f = dict(dtype=torch.float32, device=wav.device)
            H, Lc, Lx = self.conf.num_heads, self.conf.channel_layers, self.conf.cross_layers
            ret = {"logits": torch.empty((B, T, 256), **f), "vad": torch.empty((B, T, 2), **f),
                   "self_attn": torch.empty((B, 2, Lc, H, T, T), **f),
                   "cross_attn": torch.empty((B, 2, Lx, H, T, T), **f),
                   "cross_self_attn": torch.empty((B, 2, Lx, H, T, T), **f)}
            st = torch.cuda.current_stream(wav.device).cuda_stream
            _lib.check(lib, h, lib.vapb_forward_attention(h, st, wav.data_ptr(), B, S, ws.data_ptr(), ws.numel(),
                                                          *(ret[k].data_ptr() for k in ret)))
            return ret
        mode = self._mode(precision)
        ws = self._workspace(B, S, mode)
        logits = torch.empty((B, T, 256), dtype=torch.float32, device=wav.device)
        vad = torch.empty((B, T, 2), dtype=torch.float32, device=wav.device)
        st = torch.cuda.current_stream(wav.device).cuda_stream
        _lib.check(lib, h, lib.vapb_forward(h, st, wav.data_ptr(), B, S, mode, ws.data_ptr(), ws.numel(),
                                            logits.data_ptr(), vad.data_ptr()))
        return {"logits": logits, "vad": vad}

    @torch.no_grad()
    def probs(self, waveform: Tensor, vad: Optional[Tensor] = None, now_lims: List[int] = [0, 1],
              future_lims: List[int] = [2, 3], precision: Optional[str] = None,
              out: Optional[Dict[str, Tensor]] = None, counters: Optional[Tensor] = None,
              want_probs: bool = True, want_loss: bool = True) -> Dict[str, Tensor]:
        """vap/model.py:180-225. Keys: probs, vad, p_now, p_future, H, loss — `loss`
        is always present and needs T > 100, exactly like the reference (its `vad`
        argument is overwritten by the model's own sigmoid).

        Extras of the bulk path (vapb_probs_ex): `waveform` may be int16 PCM on the device (read by the fused encoder
        kernel in the 16-bit modes, converted by a library kernel otherwise); `counters` (device int64 [258]) is
        ACCUMULATED with the arg-max class histogram and the active-frame counts; want_probs / want_loss = False skip
        those two outputs (the bulk driver's compact set does not carry them)."""
        wav = self._check_input(waveform, allow_pcm16=True)
        B, _, S = wav.shape
        lib, h = _lib.load(), self._ensure_handle()
        _, T = _lib.frames(S)
        if T <= 100 and want_loss:
            raise RuntimeError(
                f"maximum size for tensor at dimension 1 is {T - 1} but size is 100"
            )
        mode = self._mode(precision)
        fmt = 0
        if wav.dtype == torch.int16:
            if mode in (_lib.MODE_BF16, _lib.MODE_FP16) and S % 2 == 0 and wav.data_ptr() % 4 == 0 and os.environ.get("VAPB_CONV01", "1") != "0":
                fmt = 1
            else:
                wav = self._pcm16_to_f32(wav)
        ws = self._workspace(B, S, mode)
        dev = wav.device
        if out is None:
            out = self.alloc_outputs(B, T, dev)
        if counters is not None and (counters.dtype != torch.int64 or counters.numel() < 258 or counters.device != dev
                                     or not counters.is_contiguous()):
            raise ValueError("counters must be a contiguous int64 tensor of 258 elements on the waveform's device")
        ptr = lambda k, want=True: out[k].data_ptr() if (want and k in out) else None
        st = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _lib.check(lib, h, lib.vapb_probs_ex(
                h, st, wav.data_ptr(), fmt, B, S, mode, ws.data_ptr(), ws.numel(),
                now_lims[0], now_lims[-1], future_lims[0], future_lims[1],
                None, None, ptr("probs", want_probs), out["vad"].data_ptr(), out["p_now"].data_ptr(),
                out["p_future"].data_ptr(), out["H"].data_ptr(), ptr("loss", want_loss), ptr("argmax"),
                None if counters is None else counters.data_ptr()))
        return out

    @staticmethod
    def alloc_outputs(B: int, T: int, device, argmax: bool = False, pin_memory: bool = False) -> Dict[str, Tensor]:
        """The six `probs()` outputs (vap/model.py:212-224); with argmax=True also the
        uint8 arg-max projection-window class per frame (an extra the bulk driver uses)."""
        f = dict(dtype=torch.float32, device=device, pin_memory=pin_memory)
        extra = {"argmax": torch.empty((B, T), dtype=torch.uint8, device=device, pin_memory=pin_memory)} if argmax else {}
        return {**{
            "probs": torch.empty((B, T, 256), **f),
            "vad": torch.empty((B, T, 2), **f),
            "p_now": torch.empty((B, T, 2), **f),
            "p_future": torch.empty((B, T, 2), **f),
            "H": torch.empty((B, T), **f),
            "loss": torch.empty((B, T - 100), **f),
        }, **extra}

    @torch.no_grad()
    def probs_host(self, waveform: Tensor, keys=("probs", "vad", "p_now", "p_future", "H", "loss"),
                   precision: Optional[str] = None, **kw) -> Dict[str, Tensor]:
        """End-to-end call with HOST buffers: what run.py does around model.probs
        (run.py:239-241): host->device copy of the waveform, the forward, and
        device->host copies of the outputs, all on the current stream."""
        if waveform.device.type != "cpu":
            raise RuntimeError("probs_host takes a CPU tensor")
        if self._device.type != "cuda":
            raise RuntimeError("call model.to('cuda') first")
        dev_wav = waveform.to(self._device, non_blocking=True)
        o = self.probs(dev_wav, precision=precision, **kw)
        host = {k: torch.empty(o[k].shape, dtype=o[k].dtype, pin_memory=True) for k in keys}
        for k in keys:
            host[k].copy_(o[k], non_blocking=True)
        torch.cuda.current_stream(self._device).synchronize()
        return host

    @torch.no_grad()
    def vad(self, waveform: Tensor, max_fill_silence_time: float = 0.02, max_omit_spike_time: float = 0.02,
            vad_cutoff: float = 0.5) -> Tensor:
        """vap/model.py:227-247: sigmoid, threshold and both run-length filters in one kernel on the model's VAD logits."""
        return self.vad_filter(self(waveform)["vad"], max_fill_silence_time, max_omit_spike_time, logits_cutoff=vad_cutoff)

    def vad_filter(self, vad01: Tensor, max_fill_silence_time: float = 0.02, max_omit_spike_time: float = 0.02,
                   logits_cutoff: Optional[float] = None) -> Tensor:
        """vad_fill_silences then vad_omit_spikes (vap/utils.py:239-272) for every item of a binary (B, T, 2)
        CUDA tensor, in place, in one kernel (vapb_vad_filter) instead of the reference's per-run Python loops.
        logits_cutoff: the tensor holds VAD logits; a frame is active when sigmoid(logit) >= logits_cutoff."""
        if vad01.device.type != "cuda":
            raise RuntimeError("vad_filter needs a CUDA tensor (no CPU fallback); utils.vad_fill_silences works on the host")
        assert vad01.ndim == 3 and vad01.shape[-1] == 2 and vad01.dtype == torch.float32 and vad01.is_contiguous()
        lib, h = _lib.load(), self._ensure_handle()
        st = torch.cuda.current_stream(vad01.device).cuda_stream
        if vad01.numel():
            _lib.check(lib, h, lib.vapb_vad_filter_ex(h, st, vad01.data_ptr(), int(logits_cutoff is not None),
                                                      0.5 if logits_cutoff is None else float(logits_cutoff),
                                                      vad01.shape[0], vad01.shape[1],
                                                      round(max_fill_silence_time * self.frame_hz),
                                                      round(max_omit_spike_time * self.frame_hz), vad01.data_ptr()))
        return vad01

    # ------------------------------------------------------------------ diagnostics
    @torch.no_grad()
    def stage(self, name: str, waveform: Tensor, precision: Optional[str] = None) -> Tensor:
        """Runs forward and returns one intermediate activation as fp32
        (vapb_get_stage): 'conv', 'ar', 'enc', 'ch', 'ar0'.., 'comb'.
        Sequence rows are channel-major (c*B + b)."""
        wav = self._check_input(waveform)
        B, _, S = wav.shape
        self.forward(wav, precision=precision)
        lib, h = _lib.load(), self._ensure_handle()
        T100, T = _lib.frames(S)
        rows = T100 if name in ("conv", "ar") else T
        nseq = B if name == "comb" else 2 * B
        out = torch.empty((nseq, rows, 256), dtype=torch.float32, device=wav.device)
        mode = self._mode(precision)
        ws = self._workspace(B, S, mode)
        st = torch.cuda.current_stream(wav.device).cuda_stream
        _lib.check(lib, h, lib.vapb_get_stage(h, st, name.encode(), B, S, mode, ws.data_ptr(), ws.numel(),
                                              out.data_ptr(), out.numel()))
        return out


VapStereo = VapGPT  # the name BASELINE.json's north_star uses for this class
