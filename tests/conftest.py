import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """Returns (recipe dict, arrays dict of torch tensors)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    recipe = ast.literal_eval(str(z["recipe"]))
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "recipe"}
    return recipe, arrays


def golden_inputs(recipe, arrays):
    """Rebuild (state_dict, waveform) of a golden case from its recipe."""
    from oracle import synth

    sd = synth.make_state_dict(recipe["seed"], recipe["ar_mode"], recipe["ar_layers"], recipe["gain"])
    if recipe.get("kind") == "example":
        wav = arrays["waveform"]
    else:
        wav = synth.make_waveform(recipe["batch"], recipe["n_samples"], recipe["wav_seed"], recipe["kind"])
    return sd, wav


CASE_NAMES = [
    "lstm1_turns_T125",
    "gru1_noise_T117",
    "lstm2_turns_T103",
    "lstm1_mono_T500",
    "lstm1_turns_T1000",
    "example_wav_T117",
]


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
