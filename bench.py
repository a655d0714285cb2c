#!/usr/bin/env python
"""Benchmark of the VAP stereo inference hot path (BASELINE.json metric:
audio-seconds processed per second on batched 20 s stereo chunks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32]
                    [--batch B] [--impl reference]

One "step" = VapGPT.probs over one batch of B synthetic 20 s stereo chunks
(configs[1]: B=256, 320 000 samples, T=1000). Prints ONE JSON line (rank 0).

 value      device-timed whole-job throughput, inputs resident in HBM.
 e2e        same metric through the bulk driver with HOST buffers: int16 PCM in
            pinned memory -> H2D -> probs -> compact outputs D2H (and, N > 1, the
            per-step all-gather / all-reduce), all inside the timed region;
            e2e.full_outputs = float32 in, all six outputs out.
 roofline   the whole step against the tensor roofline (75.6 GFLOP per chunk),
            with every kernel family timed by CUDA events on its launch stream in
            a separate profiled pass of the same steps.
 modes      every arithmetic mode (fp16 = headline, bf16, fp32): speed on the same
            batch and max-abs error against the CPU oracle on a 4-chunk sample;
            plus the oracle's torch ops run eagerly on this GPU (library bar).
 cpu_baseline  the oracle (CPU restatement of the reference, validated
            bit-identical to it) on this box's host cores, bounded sample.

--impl reference times that CPU implementation alone on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK_SAMPLES = 320_000
CHUNK_SECONDS = 20.0
METRIC = "audio-seconds per second (20 s stereo chunks, VapGPT.probs)"
UNIT = "audio-s/s"


# --------------------------------------------------------------------------- #
def flops_per_chunk(T=1000, T100=2000, lstm_layers=1, gru=False, tensor_mode=True, conv0_fused=True):
    """Algorithmic FLOPs per 20 s stereo chunk by kernel family (causal attention counted as T(T+1)/2
    pairs). Attribution follows where the work runs: in the 16-bit modes the gAR input projection runs inside the
    recurrence kernel (rnn), and with the fused encoder kernel conv0 runs inside the conv1 GEMM kernel (conv_gemm)."""
    Lc = [64000, 16000, 8000, 4000, 2000]
    if T100 != 2000:
        s = T100 / 2000.0
        Lc = [int(x * s) for x in Lc]
    conv0 = 2 * 2 * Lc[0] * 256 * 10
    conv = 2 * 2 * (Lc[1] * 256 * 2048 + (Lc[2] + Lc[3] + Lc[4]) * 256 * 1024) + 2 * 2 * T * 256 * 1280
    G = 3 if gru else 4
    rnn_in = 2 * 2 * T100 * 256 * G * 256 * lstm_layers
    rnn_rec = rnn_in
    n_att = 2 * 1 + 2 * 2 * 3  # per chunk: 2 channels x (1 self) + 2 x 3 x (self + cross)
    lin = n_att * 4 * 2 * T * 256 * 256 + 8 * 2 * 2 * T * 256 * 768 + 2 * 2 * T * 256 * 256 + 2 * T * 256 * 256
    att = n_att * 4 * 2 * 2 * 64 * (T * (T + 1) // 2)
    fused = tensor_mode and conv0_fused
    return {"conv0": 0 if fused else conv0, "conv_gemm": conv + (conv0 if fused else 0),
            "linear_gemm": lin + (0 if tensor_mode else rnn_in), "attention": att,
            "rnn": rnn_rec + (rnn_in if tensor_mode else 0), "heads": 2 * 2 * T * 256,
            "total": conv0 + conv + lin + rnn_in + att + rnn_rec}


def heads_bytes_per_chunk(T=1000):
    return 2 * T * 256 * 4 + T * 256 * 4 + T * (256 + 2 + 2 + 2 + 1) * 4 + (T - 100) * 4


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if o.returncode == 0:
                    self.rows.append([x.strip() for x in o.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm)}


# --------------------------------------------------------------------------- #
def run_reference(args, out=sys.stdout):
    """The reference's CPU implementation of the path (oracle port; the Python
    reference itself cannot travel to the GPU box), all host threads."""
    import torch

    from oracle import synth
    from oracle import vap_oracle as O

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
    b = args.ref_batch
    wav = synth.make_waveform(b, CHUNK_SAMPLES, 0, "noise")
    for _ in range(args.warmup):
        O.probs(sd, wav)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.probs(sd, wav)
    dt = time.perf_counter() - t0
    val = b * CHUNK_SECONDS * args.steps / dt
    sample = f"{b} chunks of 20 s per step ({args.steps} steps) of the B={args.batch} workload"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "fp32"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=out, flush=True)


def workload_config(args, precision):
    return {
        "workload": f"configs[1]: batched 20 s stereo chunks, B={args.batch} per GPU, 320000 samples, T=1000",
        "batch_per_gpu": args.batch, "precision": precision,
        "l2_policy": "inputs larger than L2 (waveform batch %.0f MB, activations GBs)" % (args.batch * 2.56),
        "weights": "synthetic seed 0 (reference schema, LSTMx1); shipped checkpoints are absent",
        "audio": getattr(args, "audio", "noise"),
    }


def cpu_baseline(sample_batch=2, iters=2):
    import torch

    from oracle import synth
    from oracle import vap_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
    wav = synth.make_waveform(sample_batch, CHUNK_SAMPLES, 0, "noise")
    O.probs(sd, wav)
    t0 = time.perf_counter()
    for _ in range(iters):
        O.probs(sd, wav)
    dt = time.perf_counter() - t0
    return {"value": sample_batch * CHUNK_SECONDS * iters / dt, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port", "sample": f"{iters} x {sample_batch} chunks of 20 s (oracle port of VapGPT.probs, fp32)"}


# --------------------------------------------------------------------------- #
def _bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs local to its GPU (NVML affinity mask) BEFORE pinned buffers are allocated, so
    the host side of the H2D / D2H copies is first-touched on the GPU's own NUMA node (matters when 8 ranks stream
    ~1 GB per step each through the same host)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def main():
    # stdout carries exactly one JSON line: anything a library prints to fd 1 (NCCL's version banner, ...) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--precision", default=os.environ.get("VAPB_BENCH_PRECISION", "auto"))
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ref-batch", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the per-mode speed/error table and the eager bar")
    ap.add_argument("--audio", default="noise", choices=["noise", "turns"],
                    help="synthetic input: 0.05*N(0,1), or SURVEY §8d config 2's turn-taking variant (gated noise + tone)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args, real_stdout)
        return

    import torch

    from oracle import synth
    from voiceactivityprojection_b200 import VapConfig, VapGPT, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    local_cpus = _bind_to_gpu_numa_node(local) if world > 1 else 0

    sd = synth.make_state_dict(0, "LSTM", 1, 2.0)
    model = VapGPT(VapConfig()).to(dev)
    model.load_state_dict(sd)
    precision = args.precision
    B = args.batch
    # synthetic audio generated on the device, seeded by global chunk ids
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    wav = torch.randn((B, 2, CHUNK_SAMPLES), generator=gen, device=dev, dtype=torch.float32) * 0.05
    if args.audio == "turns":
        wav = synth.make_waveform(B, CHUNK_SAMPLES, 1 + rank, "turns").to(dev)
    if precision == "auto":
        # the headline mode is the tensor-core mode that meets the stated tolerance (probs <= 1e-3 of the reference's
        # fp32 outputs): fp16 operands. bf16 (same kernels, same speed, 8x the error) and fp32 are reported in `modes`.
        precision = "fp16"
    T = 1000
    out = model.alloc_outputs(B, T, dev)

    def timed_steps(prec, n_steps, n_warm):
        """Device-timed probs() steps of one mode on the resident batch (CUDA events on the launch stream)."""
        for _ in range(n_warm):
            model.probs(wav, precision=prec, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_steps):
            model.probs(wav, precision=prec, out=out)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n_steps

    def step():
        model.probs(wav, precision=precision, out=out)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = model.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = model.launch_count() - l0
    if dist:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        cnt = torch.tensor([float(out["p_now"].shape[0])], device=dev)  # chunks this rank processed per step
        dist.all_reduce(cnt)
        total_chunks = int(cnt.item())
    else:
        total_chunks = B
    value = total_chunks * CHUNK_SECONDS * args.steps / (ms / 1e3)

    # ---- end to end through the bulk driver (the call a bulk-inference user makes; BASELINE configs[3]). Default
    # variant = what SURVEY §8e/§8f specify: every step's batch starts as int16 PCM in pinned HOST memory (what wav
    # files hold), the compact per-chunk outputs (vad, p_now, p_future, H, arg-max class: 29 KB per chunk) end in pinned
    # HOST memory, and at N > 1 every step's all-gather of the compact buffer and all-reduce of the step's counters run
    # on a side stream INSIDE the timed region. `full_outputs` = the reference's float32 waveforms in and all six
    # probs() outputs out (1.05 MB per chunk). H2D of step i+1, the forward of step i, the collectives and the D2H of
    # step i-1 overlap on four streams (voiceactivityprojection_b200/bulk.py). Copies are inside the timed region.
    e2e = None
    if not args.no_e2e:
        from voiceactivityprojection_b200.bulk import ALL_KEYS, COMPACT_KEYS, BulkRunner

        seen = []
        n_e2e = max(8, 2 * args.steps)  # the pipeline's fill (first H2D) and drain (last D2H) are inside the timed region

        def run_e2e(host_batch, **kw):
            runner = BulkRunner(model, B, CHUNK_SAMPLES, precision=precision, stats=True, **kw)
            runner.run([host_batch] * 2, sink=lambda i, b, o: seen.append(float(o["p_now"][0, 0, 0])))
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            runner.h2d_bytes = runner.d2h_bytes = runner.coll_bytes = 0
            t0 = time.perf_counter()
            stats = runner.run((host_batch for _ in range(n_e2e)),
                               sink=lambda i, b, o: seen.append(float(o["p_now"][0, 0, 0])))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist:
                t = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            res = {"value": total_chunks * CHUNK_SECONDS * n_e2e / dt, "unit": UNIT,
                   "h2d_bytes_per_step": runner.h2d_bytes // n_e2e, "d2h_bytes_per_step": runner.d2h_bytes // n_e2e,
                   "steps": n_e2e}
            if runner.gather:
                res["collective_bytes_per_step"] = runner.coll_bytes // n_e2e
            return res, stats, runner

        host_pcm = (wav * 32768.0).clamp_(-32768, 32767).to(torch.int16).cpu().pin_memory()
        e2e, stats, runner = run_e2e(host_pcm, keys=COMPACT_KEYS, gather=bool(dist))
        e2e["api"] = ("BulkRunner.run: int16 PCM pinned host batches in (read by the encoder kernel), compact outputs "
                      "(vad, p_now, p_future, H, argmax) to pinned host memory, 4-stream pipeline")
        e2e["bulk"] = {"chunks": stats.chunks, "frames": stats.frames, "vad_active_frames": stats.vad_active.tolist(),
                       "classes_seen": int((stats.class_hist > 0).sum()),
                       "compact_bytes_per_chunk": runner.layout.bytes_per_chunk,
                       "counters": "arg-max class histogram + active frames, taken by the heads kernel"}
        if dist:
            e2e["bulk"]["collectives"] = ("per step, timed: ncclAllGather(compact buffer, %d B per rank) + "
                                          "ncclAllReduce(258 counters) on a side stream; rank 0 copies the gathered set "
                                          "to the host" % runner.layout.nbytes)
            # per-step counters were all-reduced on the device: totals are global
            assert int(stats.class_hist.sum()) == world * B * T * n_e2e, (int(stats.class_hist.sum()), world * B * T * n_e2e)
        else:
            assert int(stats.class_hist.sum()) == B * T * n_e2e
        del runner, host_pcm
        host_wav = torch.empty((B, 2, CHUNK_SAMPLES), dtype=torch.float32, pin_memory=True)
        host_wav.copy_(wav)
        full, _, runner = run_e2e(host_wav, keys=ALL_KEYS)
        full["api"] = "float32 waveforms in, all six probs() outputs out (the reference's interface through the same pipeline)"
        e2e["full_outputs"] = full
        del runner, host_wav
    clocks = sampler.stop() if sampler else None  # sampled over both timed regions (device-timed steps and e2e)

    # ---- roofline: profiled pass of the same steps (per-family CUDA-event timing on the launch stream)
    roofline = None
    modes = None
    if rank == 0:
        import ctypes as C

        lib, h = _lib.load(), model._ensure_handle()
        lib.vapb_profile_begin(h)
        n_prof = min(args.steps, 3)
        for _ in range(n_prof):
            step()
        fam_ms = (C.c_double * 7)()
        fam_n = (C.c_uint64 * 7)()
        _lib.check(lib, h, lib.vapb_profile_end(h, fam_ms, fam_n))
        fams = {n: (fam_ms[i] / n_prof, int(fam_n[i]) // n_prof) for i, n in enumerate(_lib.PROFILE_FAMILIES)}
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        tensor_mode = precision in ("bf16", "fp16")  # fp32_tc keeps the fp32 path's kernel split (conv0, xproj, ...)
        fl = flops_per_chunk(tensor_mode=tensor_mode, conv0_fused=fams["conv0"][0] == 0.0)
        dom = max(("conv_gemm", "linear_gemm", "attention", "rnn", "conv0"), key=lambda k: fams[k][0])
        dom_ms, dom_n = fams[dom]
        step_ms = sum(v[0] for v in fams.values())
        if tensor_mode:
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
        else:
            peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
            peak_src = "fp32 CUDA-core FMA peak 148 SM x 128 FMA x 2 x sm_max_mhz (fp32 mode does not use the tensor pipe)"
        tfl = lambda f, t_ms: (fl[f] * B / (t_ms / 1e3) / 1e12) if t_ms > 0 else None
        whole = fl["total"] * total_chunks * args.steps / (ms / 1e3) / 1e12 / world
        # DRAM bytes per step from the newest committed ncu capture of this command (not measured in this run)
        traffic, traffic_src = None, None
        for name in ("r2_traffic_fp16.json", "r1_traffic_bf16.json"):
            tf = os.path.join(ROOT, "profiles", name)
            if tensor_mode and B == 256 and os.path.exists(tf):
                tj = json.load(open(tf))
                traffic = tj.get("total", {}).get("dram_bytes_per_step")
                traffic_src = "committed ncu capture profiles/%s (not measured in this run): %s" % (name, tj.get("_source", ""))
                break
        roofline = {
            # the judged figure: the WHOLE step against the tensor roofline (75.6 GFLOP per chunk)
            "kernel": "whole step (all families)", "bound": "tensor", "achieved": whole, "peak": peak, "unit": "TFLOP/s",
            "frac": whole / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src, "algorithmic_gflop_per_chunk": fl["total"] / 1e9,
            "families_ms_per_step": {k: round(v[0], 3) for k, v in fams.items()},
            "families_tflops": {k: (round(tfl(k, v[0]), 1) if tfl(k, v[0]) else None) for k, v in fams.items()
                                if k in ("conv0", "conv_gemm", "linear_gemm", "attention", "rnn")},
            "families_frac_of_peak": {k: (round(tfl(k, v[0]) / peak, 3) if tfl(k, v[0]) else None) for k, v in fams.items()
                                      if k in ("conv0", "conv_gemm", "linear_gemm", "attention", "rnn")},
            "dominant": {"kernel": dom, "ms_per_step": dom_ms, "launches_per_step": dom_n,
                         "share_of_step": dom_ms / step_ms if step_ms else None, "achieved": tfl(dom, dom_ms),
                         "frac": (tfl(dom, dom_ms) / peak) if tfl(dom, dom_ms) else None},
            "attribution": "gAR input projection counted under rnn (it runs inside the recurrence kernel); conv0 under "
                           "conv_gemm when the fused encoder kernel is on" if tensor_mode else "fp32 path: separate kernels",
            "heads": {"bound": "hbm", "achieved": heads_bytes_per_chunk() * B / (fams["heads"][0] / 1e3) / 1e9 if fams["heads"][0] else None,
                      "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s"},
        }

        # ---- every arithmetic mode, same batch: speed, and error against the CPU oracle (the reference's fp32
        # forward, restated) on a 4-chunk turn-taking sample. The headline `value` is modes[precision].
        if not args.no_modes:
            from oracle import vap_oracle as O

            sw = synth.make_waveform(4, CHUNK_SAMPLES, 11, "turns")
            torch.set_num_threads(os.cpu_count() or 1)
            ref = O.probs(sd, sw)
            ref_fwd = O.forward(sd, sw)
            modes = {}
            for prec, (n_steps, n_warm) in (("fp16", (4, 2)), ("bf16", (4, 2)), ("fp32_tc", (3, 1)), ("fp32", (2, 1))):
                t_ms = timed_steps(prec, n_steps, n_warm)
                o = model.probs(sw.to(dev), precision=prec)
                f = model.forward(sw.to(dev), precision=prec)
                err = {k: float((o[k].cpu() - ref[k]).abs().max()) for k in ("probs", "vad", "p_now", "p_future", "H")}
                err["logits"] = float((f["logits"].cpu() - ref_fwd["logits"]).abs().max())
                agree = float((o["probs"].argmax(-1).cpu() == ref["probs"].argmax(-1)).float().mean())
                vad_same = float(((o["vad"].cpu() >= 0.5) == (ref["vad"] >= 0.5)).float().mean())
                modes[prec] = {"value": B * CHUNK_SECONDS / (t_ms / 1e3), "unit": UNIT, "ms_per_step": t_ms,
                               "steps": n_steps, "max_abs_err_vs_oracle": err, "argmax_agreement": agree,
                               "vad_threshold_agreement": vad_same}
            modes["_sample"] = "4 chunks of 20 s, turn-taking synthetic audio (seed 11), synthetic weights seed 0"
            # the library-kernel bar: the oracle's torch ops (cuDNN conv/LSTM, cuBLAS, eager softmax) on this same GPU
            try:
                sd_dev = {k: v.to(dev) for k, v in sd.items()}
                eb = min(B, 32)
                with torch.no_grad():
                    for name, tf32, ac in (("fp32", False, False), ("tf32", True, False), ("bf16_autocast", True, True)):
                        torch.backends.cuda.matmul.allow_tf32 = tf32
                        torch.backends.cudnn.allow_tf32 = tf32
                        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                            O.probs(sd_dev, wav[:eb])
                            torch.cuda.synchronize()
                            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            a.record()
                            O.probs(sd_dev, wav[:eb])
                            b.record()
                            torch.cuda.synchronize()
                        modes.setdefault("eager_b200", {})[name] = {
                            "value": eb * CHUNK_SECONDS / (a.elapsed_time(b) / 1e3), "unit": UNIT, "batch": eb}
                modes["eager_b200"]["_what"] = ("oracle restatement (the reference's ATen ops) run with torch eager on "
                                                "this GPU: the library-kernel bar, not the product path")
                del sd_dev
            except Exception as e:  # an out-of-memory eager run must not take the bench line with it
                modes["eager_b200"] = {"unavailable": repr(e)[:200]}
            torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "fp32_tc": "f16x2 split (fp32-class)"}.get(precision, "f32"),
            "data": "synthetic", "config": workload_config(args, precision), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "modes": modes,
        }
        if local_cpus:
            line["config"]["host_affinity"] = f"rank bound to the {local_cpus} CPUs local to its GPU (NVML)"
        print(json.dumps(line), file=real_stdout, flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
