"""Writes tests/golden/vad_list.json with the UNMODIFIED reference's voice-activity list helpers
(vap/utils.py: get_vad_list_subset :141-167, vad_list_to_onehot :170-195, vad_onehot_to_vad_list :198-236,
get_dialog_states :130-138) on seeded cases, including segments that coincide with the window boundaries, and on
the reference's own example/student_long_female_en-US-Wavenet-G_vad_list.json. TEST INFRASTRUCTURE.
Run in the authoring container: python oracle/make_golden_vadlist.py"""
import json
import os
import random
import sys

import torch

sys.path.insert(0, "/root/reference")
from vap import utils as R  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
random.seed(0)
out = {"subset": [], "onehot_to_list": [], "list_to_onehot": []}
example = json.load(open("/root/reference/example/student_long_female_en-US-Wavenet-G_vad_list.json"))
out["example_vad_list"] = example
for a, b in ((0.0, 1.0), (1.0, 3.0), (2.12, 4.0), (0.5, 2.12)):
    out["subset"].append({"vad_list": example, "start": a, "end": b, "out": R.get_vad_list_subset(example, a, b)})
for _ in range(40):
    lists = []
    for ch in range(2):
        t, segs = 0.0, []
        for _ in range(random.randint(0, 6)):
            t += random.choice([0, 0.25, 0.5, 1.0])
            s0 = t
            t += random.choice([0.25, 0.5, 1.0, 2.0])
            segs.append([s0, t])
        lists.append(segs)
    a = random.choice([0, 0.5, 1.0, 2.0, 3.0])
    b = a + random.choice([0.5, 1.0, 2.0, 4.0])
    out["subset"].append({"vad_list": lists, "start": a, "end": b, "out": R.get_vad_list_subset(lists, a, b)})
g = torch.Generator().manual_seed(0)
for thr in (0.1, 0.0, 0.3):
    flips = (torch.rand((2, 300, 2), generator=g) < 0.06).long()
    v = (flips.cumsum(1) % 2).float()
    v[0, :, 1] = 0.0
    lists = R.vad_onehot_to_vad_list(v, 50, thr)
    out["onehot_to_list"].append({"vad": v.long().tolist(), "frame_hz": 50, "ipu_thresh_time": thr, "out": lists})
    for kw in (dict(frame_hz=50), dict(hop_time=0.02), dict(frame_hz=100, channel_first=True)):
        oh = R.vad_list_to_onehot(lists[1], 6.0, **kw)
        out["list_to_onehot"].append({"vad_list": lists[1], "duration": 6.0, "kw": kw, "out": oh.long().tolist()})
oh = R.vad_list_to_onehot(example, 12.0, frame_hz=50)
out["list_to_onehot"].append({"vad_list": example, "duration": 12.0, "kw": {"frame_hz": 50}, "out": oh.long().tolist()})
v = torch.randint(0, 2, (3, 7, 2), generator=g).float()
out["dialog_states"] = {"vad": v.long().tolist(), "out": R.get_dialog_states(v).tolist()}
path = os.path.join(ROOT, "tests", "golden", "vad_list.json")
json.dump(out, open(path, "w"))
print(path, os.path.getsize(path), "bytes;", {k: len(v) for k, v in out.items() if isinstance(v, list)})
