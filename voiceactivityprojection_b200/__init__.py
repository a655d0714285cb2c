"""B200-native implementation of the VAP stereo inference forward path
(drop-in for the reference's `vap.model.VapGPT` forward / probs / vad)."""
from .model import VapConfig, VapGPT, VapStereo, load_older_state_dict  # noqa: F401
from .objective import ObjectiveVAP  # noqa: F401
