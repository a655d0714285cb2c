"""GPU parity of the device resampler (vapb_resample through audio.resample_device) against the numpy oracle and
the torchaudio golden vectors. Tolerance: float32 sums of <= 475 products of |x| <= 1 samples in a different order
than torchaudio's conv1d: max-abs 2e-6 (measured ~3e-7)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
TOL = 2e-6


def _gold():
    return np.load(f"{GOLDEN_DIR}/resample_example_24k_16k.npz")


def test_example_wav_pcm_24k_to_16k_matches_torchaudio_golden():
    from voiceactivityprojection_b200.audio import resample_device

    g = _gold()
    pcm = torch.from_numpy(g["pcm_24k"]).cuda()
    y = resample_device(pcm, 24000, 16000)                      # int16 in, scaled on the device
    assert y.shape == (24000,) and y.dtype == torch.float32
    assert np.abs(y.cpu().numpy() - g["y_24k"]).max() <= TOL
    yf = resample_device(pcm.float() / 32768.0, 24000, 16000)   # float in
    assert torch.equal(y, yf)


@pytest.mark.parametrize("rate", [48000, 44100, 8000])
def test_other_rates_match_torchaudio_golden_and_oracle(rate):
    from oracle import resample_oracle as R
    from voiceactivityprojection_b200.audio import resample_device

    g = _gold()
    x = g[f"x_{rate}"]
    y = resample_device(torch.from_numpy(x).cuda(), rate, 16000).cpu().numpy()
    assert y.shape == g[f"y_{rate}"].shape
    assert np.abs(y - g[f"y_{rate}"]).max() <= TOL
    assert np.abs(y - R.resample(x, rate, 16000)).max() <= TOL


def test_interleaved_stereo_pcm_ragged_lengths_and_empty():
    from oracle import resample_oracle as R
    from voiceactivityprojection_b200.audio import resample_device

    rng = np.random.default_rng(3)
    for items, n in ((1, 1), (3, 2), (2, 7), (5, 2999), (4, 48000)):
        pcm = rng.integers(-32768, 32767, size=(items, n, 2), dtype=np.int16)
        y = resample_device(torch.from_numpy(pcm).cuda(), 24000, 16000, interleaved=True)
        ref = R.resample(np.ascontiguousarray(pcm.transpose(0, 2, 1)), 24000, 16000)
        assert y.shape == ref.shape == (items, 2, -(-2 * n // 3))
        assert np.abs(y.cpu().numpy() - ref).max() <= TOL
    e = resample_device(torch.empty((2, 0), dtype=torch.float32, device="cuda"), 24000, 16000)
    assert e.shape == (2, 0)
    with pytest.raises(RuntimeError):
        resample_device(torch.zeros(4, 10), 24000, 16000)  # CPU tensor: no fallback
    with pytest.raises(TypeError):
        resample_device(torch.zeros(4, 10, dtype=torch.float64, device="cuda"), 24000, 16000)


def test_bulk_runner_ingests_24k_pcm_and_matches_host_resampled_input():
    """BulkRunner(input_rate=24000, pcm16=True): 24 kHz int16 crosses PCIe (0.75x the bytes of 16 kHz float32... per
    second of audio: 96 KB vs 128 KB) and is resampled on the device; outputs equal those of feeding the model the
    device-resampled waveform directly, and are within fp32 tolerance of the host (oracle) resampled input."""
    from oracle import resample_oracle as R, synth
    from voiceactivityprojection_b200 import VapConfig, VapGPT
    from voiceactivityprojection_b200.audio import resample_device
    from voiceactivityprojection_b200.bulk import BulkRunner

    m = VapGPT(VapConfig(), precision="fp32").to("cuda")
    m.load_state_dict(synth.make_state_dict(7, "LSTM", 1, 2.0))
    n16 = 48000
    runner = BulkRunner(m, batch=3, n_samples=n16, pcm16=True, input_rate=24000, stats=False,
                        keys=("probs", "vad", "p_now", "p_future", "H"))
    assert runner.n_in == 72000
    rng = np.random.default_rng(4)
    pcm = torch.from_numpy((rng.standard_normal((5, 2, 72000)) * 1500).astype(np.int16))
    got = {}
    runner.run([pcm[:3].pin_memory(), pcm[3:].pin_memory()],
               sink=lambda i, b, o: got.update({i: {k: v[:b].clone() for k, v in o.items()}}))
    wav = resample_device(pcm.cuda(), 24000, 16000)
    assert wav.shape == (5, 2, n16)
    direct = m.probs(wav)
    host = m.probs(torch.from_numpy(R.resample(pcm.numpy(), 24000, 16000)).cuda())
    for k in ("probs", "vad", "p_now", "p_future", "H"):
        both = torch.cat([got[0][k], got[1][k]])
        assert torch.equal(both, direct[k].cpu()), k
        assert (both - host[k].cpu()).abs().max().item() <= 1e-4, k


def test_load_waveform_on_device_resamples_with_the_device_kernel(tmp_path):
    from scipy.io import wavfile

    from voiceactivityprojection_b200.audio import load_waveform

    g = _gold()
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 24000, np.stack([g["pcm_24k"], g["pcm_24k"][::-1]], axis=1))
    host, sr = load_waveform(p, sample_rate=16000)                 # torchaudio on the CPU, as the reference
    dev, sr2 = load_waveform(p, sample_rate=16000, device="cuda")  # vapb_resample
    assert sr == sr2 == 16000 and dev.is_cuda and dev.shape == host.shape == (2, 24000)
    assert (dev.cpu() - host).abs().max().item() <= TOL
