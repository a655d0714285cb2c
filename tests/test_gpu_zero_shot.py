"""GPU parity of the ZeroShot marginals kernel (vapb_zero_shot through voiceactivityprojection_b200.zero_shot)
against the unmodified reference's outputs (tests/golden/zero_shot.npz, oracle/make_golden_zeroshot.py) and the
numpy oracle. Tolerance: float32 softmax + sums of <= 56 probabilities in a different order + one division:
max-abs 2e-6 on values in [0, 1]."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu
TOL = 2e-6


def _zs():
    from voiceactivityprojection_b200.zero_shot import ZeroShot

    return ZeroShot(bin_times=[0.2, 0.4, 0.6, 0.8], frame_hz=50)


@pytest.mark.parametrize("name", ["flat", "peaked", "odd"])
def test_get_probs_matches_reference_golden(name):
    g = np.load(f"{GOLDEN_DIR}/zero_shot.npz")
    zs = _zs()
    logits = torch.from_numpy(g[name + "_logits"]).cuda()
    va = torch.from_numpy(g[name + "_va"]).cuda()
    out = zs.get_probs(logits, va)
    assert set(out) == {"p", "p_bc"}
    for k in ("p", "p_bc"):
        assert out[k].shape == g[f"{name}_{k}"].shape and out[k].dtype == torch.float32
        assert np.abs(out[k].cpu().numpy() - g[f"{name}_{k}"]).max() <= TOL, k
    probs = logits.softmax(-1)
    for k, fn in (("p_sil", zs.probs_on_silence), ("p_act", zs.probs_on_active), ("p_bc", zs.probs_backchannel)):
        assert np.abs(fn(probs).cpu().numpy() - g[f"{name}_{k}"]).max() <= TOL, k
    p2 = zs.probs_next_speaker(probs, va)
    assert np.abs(p2.cpu().numpy() - g[name + "_p"]).max() <= TOL


def test_full_size_properties_and_oracle_sample():
    """B=256 x T=1000 (BASELINE configs[1]): p is a distribution over the two speakers in every dialog state,
    the silence marginals are complementary, p_bc is a probability; 512 sampled frames equal the oracle."""
    from oracle import zero_shot_oracle as Z

    zs = _zs()
    g = torch.Generator(device="cuda").manual_seed(3)
    B, T = 256, 1000
    logits = torch.randn((B, T, 256), generator=g, device="cuda") * 4.0
    va = (torch.rand((B, T, 2), generator=g, device="cuda") < 0.5).float()
    out = zs.get_probs(logits, va)
    p, p_bc = out["p"], out["p_bc"]
    assert torch.isfinite(p).all() and torch.isfinite(p_bc).all()
    assert (p.sum(-1) - 1).abs().max().item() <= 1e-6
    assert p.min().item() >= 0 and p_bc.min().item() >= 0 and p_bc.sum(-1).max().item() <= 1 + 1e-6
    sil = zs.probs_on_silence(logits.softmax(-1))
    assert (sil.sum(-1) - 1).abs().max().item() <= 1e-6
    idx = torch.randint(0, B * T, (512,), generator=torch.Generator().manual_seed(4))
    lg = logits.view(-1, 256)[idx.cuda()].cpu().numpy()[None]
    v = va.view(-1, 2)[idx.cuda()].cpu().numpy()[None]
    ref = Z.get_probs(lg, v)
    assert np.abs(p.view(-1, 2)[idx.cuda()].cpu().numpy() - ref["p"][0]).max() <= TOL
    assert np.abs(p_bc.view(-1, 2)[idx.cuda()].cpu().numpy() - ref["p_bc"][0]).max() <= TOL


def test_dialog_states_other_than_the_four_give_zero_and_short_va_is_refused():
    from voiceactivityprojection_b200 import _lib

    zs = _zs()
    logits = torch.randn((1, 6, 256), device="cuda")
    va = torch.tensor([[[0., 0.], [1., 0.], [1., 1.], [0., 1.], [3., 0.], [0., 2.]]], device="cuda")
    p = zs.get_probs(logits, va)["p"][0].cpu()
    # (3, 0) -> state -2, (0, 2) -> state 5: no branch of probs_next_speaker writes them (zero_shot.py:239-262)
    assert torch.equal(p[4:], torch.zeros(2, 2)) and (p[:4].sum(-1) - 1).abs().max() <= 1e-6
    with pytest.raises(AssertionError):
        zs.get_probs(logits, va[:, :5])
    lib = _lib.load()
    buf = torch.empty((1, 6, 2), device="cuda")
    rc = lib.vapb_zero_shot(None, None, logits.data_ptr(), 0, 1, 6, va.data_ptr(), 5, zs._sets, buf.data_ptr(), None,
                            None, None)
    assert rc != 0 and b"voice activity" in lib.vapb_last_error(None)


def test_empty_batch_and_model_owned_objective_counts_launches():
    from oracle import synth
    from voiceactivityprojection_b200 import VapConfig, VapGPT
    from voiceactivityprojection_b200.zero_shot import ZeroShot

    zs = _zs()
    out = zs.get_probs(torch.zeros((0, 5, 256), device="cuda"), torch.zeros((0, 5, 2), device="cuda"))
    assert out["p"].shape == (0, 5, 2)
    model = VapGPT(VapConfig()).to("cuda:0")
    model.load_state_dict(synth.make_state_dict(1, "LSTM", 1, 1.0))
    zs2 = ZeroShot(bin_times=model.conf.bin_times, frame_hz=model.frame_hz)
    zs2._owner = model
    n0 = model.launch_count()
    zs2.get_probs(torch.randn((2, 9, 256), device="cuda"), torch.zeros((2, 9, 2), device="cuda"))
    assert model.launch_count() == n0 + 1
