#!/usr/bin/env python
"""Benchmark of the VAP stereo inference hot path (BASELINE.json metric:
audio-seconds processed per second on batched 20 s stereo chunks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32]
                    [--batch B] [--impl reference]

One "step" = VapGPT.probs over one batch of B synthetic 20 s stereo chunks
(configs[1]: B=256, 320 000 samples, T=1000). Prints ONE JSON line (rank 0).

 value      device-timed whole-job throughput, inputs resident in HBM.
 e2e        same metric through the public host-buffer call (pinned host
            waveform -> H2D -> probs -> D2H of every output), copies timed.
 roofline   dominant kernel family, timed with CUDA events on its launch stream
            in a separate profiled pass of the same steps.
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
def flops_per_chunk(T=1000, T100=2000, lstm_layers=1, gru=False):
    """Algorithmic FLOPs per 20 s stereo chunk by kernel family (SURVEY.md §8d:
    causal attention counted as T(T+1)/2 pairs)."""
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
    return {"conv0": conv0, "conv_gemm": conv, "linear_gemm": lin + rnn_in, "attention": att, "rnn": rnn_rec,
            "heads": 2 * 2 * T * 256, "total": conv0 + conv + lin + rnn_in + att + rnn_rec}


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

    model = VapGPT(VapConfig()).to(dev)
    model.load_state_dict(synth.make_state_dict(0, "LSTM", 1, 2.0))
    precision = args.precision
    B = args.batch
    # synthetic audio generated on the device, seeded by global chunk ids
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    wav = torch.randn((B, 2, CHUNK_SAMPLES), generator=gen, device=dev, dtype=torch.float32) * 0.05
    if args.audio == "turns":
        wav = synth.make_waveform(B, CHUNK_SAMPLES, 1 + rank, "turns").to(dev)
    if precision == "auto":
        try:
            model.probs(wav[:1], precision="bf16")
            precision = "bf16"
        except Exception:
            precision = "fp32"
    T = 1000
    out = model.alloc_outputs(B, T, dev)

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

    # ---- end to end through the bulk driver (the call a bulk-inference user makes): every step's batch starts in
    # pinned HOST memory and every output ends in pinned HOST memory; H2D of step i+1, the forward of step i and the
    # D2H of step i-1 overlap on three streams (voiceactivityprojection_b200/bulk.py). Copies are inside the timed region.
    e2e = None
    if not args.no_e2e:
        from voiceactivityprojection_b200.bulk import ALL_KEYS, BulkRunner

        host_wav = torch.empty((B, 2, CHUNK_SAMPLES), dtype=torch.float32, pin_memory=True)
        host_wav.copy_(wav)
        runner = BulkRunner(model, B, CHUNK_SAMPLES, precision=precision, keys=ALL_KEYS, stats=True)
        seen = []
        runner.run([host_wav] * 2, sink=lambda i, b, o: seen.append(float(o["p_now"][0, 0, 0])))
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        n_e2e = max(8, 2 * args.steps)  # the pipeline's fill (first H2D) and drain (last D2H) are inside the timed region
        runner.h2d_bytes = runner.d2h_bytes = 0
        t0 = time.perf_counter()
        stats = runner.run((host_wav for _ in range(n_e2e)), sink=lambda i, b, o: seen.append(float(o["p_now"][0, 0, 0])))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total_chunks * CHUNK_SECONDS * n_e2e / dt, "unit": UNIT,
               "h2d_bytes_per_step": runner.h2d_bytes // n_e2e, "d2h_bytes_per_step": runner.d2h_bytes // n_e2e,
               "steps": n_e2e, "api": "BulkRunner.run (pinned host batches in, pinned host outputs out, 3-stream pipeline)"}
        # the only collectives of the path (BASELINE configs[3]): all-reduce of the shard counters and all-gather of the
        # compact per-chunk outputs of the last batch, over NCCL on the devices (outside the timed region)
        bulk = None
        if dist:
            from voiceactivityprojection_b200.bulk import gather_compact

            tot = stats.all_reduce(device=dev)
            last = runner.dout[(n_e2e - 1) % runner.depth]
            comp = gather_compact({k: last[k] for k in ("vad", "p_now", "p_future", "H", "argmax")})
            bulk = {"chunks": tot.chunks, "frames": tot.frames, "vad_active_frames": tot.vad_active.tolist(),
                    "classes_seen": int((tot.class_hist > 0).sum()), "gathered_chunks": int(comp["p_now"].shape[0]),
                    "gathered_bytes": int(sum(v.numel() * v.element_size() for v in comp.values())),
                    "collectives": "ncclAllReduce(counters) + ncclAllGather(compact outputs)"}
            assert tot.chunks == world * B * n_e2e and comp["p_now"].shape[0] == world * B
            e2e["bulk"] = bulk
        del runner
        # the same pipeline fed with the int16 PCM a wav file holds (BulkRunner(pcm16=True): half the H2D bytes, scaled
        # on the device) — reported beside `e2e`, which keeps the reference's float32 waveforms; it matters when
        # several GPUs share the host's PCIe uplinks
        host_pcm = (host_wav * 32768.0).clamp_(-32768, 32767).to(torch.int16).pin_memory()
        runner = BulkRunner(model, B, CHUNK_SAMPLES, precision=precision, keys=ALL_KEYS, stats=True, pcm16=True)
        runner.run([host_pcm] * 2)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        runner.h2d_bytes = runner.d2h_bytes = 0
        t0 = time.perf_counter()
        runner.run((host_pcm for _ in range(n_e2e)), sink=lambda i, b, o: seen.append(float(o["p_now"][0, 0, 0])))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if dist:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e["pcm16_input"] = {"value": total_chunks * CHUNK_SECONDS * n_e2e / dt, "unit": UNIT,
                              "h2d_bytes_per_step": runner.h2d_bytes // n_e2e,
                              "d2h_bytes_per_step": runner.d2h_bytes // n_e2e, "steps": n_e2e}
        del runner, host_pcm
    clocks = sampler.stop() if sampler else None  # sampled over both timed regions (device-timed steps and e2e)

    # ---- roofline of the dominant kernel family: profiled pass of the same steps
    roofline = None
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
        fl = flops_per_chunk()
        dom = max(("conv_gemm", "linear_gemm", "attention", "rnn", "conv0"), key=lambda k: fams[k][0])
        dom_ms, dom_n = fams[dom]
        step_ms = sum(v[0] for v in fams.values())
        if precision in ("bf16", "fp16"):
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
        else:
            peak = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
            peak_src = "fp32 CUDA-core FMA peak 148 SM x 128 FMA x 2 x sm_max_mhz (fp32 mode does not use the tensor pipe)"
        ach = fl[dom] * B / (dom_ms / 1e3) / 1e12 if dom_ms > 0 else 0.0
        # DRAM bytes of the dominant family per step, from the committed ncu capture of this same command
        traffic, traffic_src = None, None
        tf = os.path.join(ROOT, "profiles", "r1_traffic_bf16.json")
        if precision in ("bf16", "fp16") and B == 256 and os.path.exists(tf):
            tj = json.load(open(tf))
            if dom in tj:
                traffic, traffic_src = tj[dom]["dram_bytes_per_step"], "profiles/r1_traffic_bf16.json: " + tj["_source"]
        roofline = {
            "kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": ach / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "ms_per_step": dom_ms, "launches_per_step": dom_n, "share_of_step": dom_ms / step_ms if step_ms else None,
            "families_ms_per_step": {k: round(v[0], 3) for k, v in fams.items()},
            "whole_step": {"achieved": fl["total"] * total_chunks * args.steps / (ms / 1e3) / 1e12 / world,
                           "unit": "TFLOP/s per GPU", "frac": fl["total"] * total_chunks * args.steps / (ms / 1e3) / 1e12 / world / peak},
            "heads": {"bound": "hbm", "achieved": heads_bytes_per_chunk() * B / (fams["heads"][0] / 1e3) / 1e9 if fams["heads"][0] else None,
                      "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s"},
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "f16"}.get(precision, "f32"),
            "data": "synthetic", "config": workload_config(args, precision), "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        }
        if local_cpus:
            line["config"]["host_affinity"] = f"rank bound to the {local_cpus} CPUs local to its GPU (NVML)"
        print(json.dumps(line), file=real_stdout, flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
