"""B200-native implementation of the VAP stereo inference forward path
(drop-in for the reference's `vap.model.VapGPT` forward / probs / vad)."""
from .model import VapConfig, VapGPT, VapStereo, load_older_state_dict  # noqa: F401
from .objective import ObjectiveVAP  # noqa: F401
from .zero_shot import ZeroShot  # noqa: F401
from .bulk import BulkRunner  # noqa: F401,E402
from .session import step_extraction  # noqa: F401,E402
from .streaming import StreamingVAP  # noqa: F401,E402
