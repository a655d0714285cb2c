"""Inference CLI with the reference's `run.py` contract (run.py:134-262):

    python -m voiceactivityprojection_b200.run --audio X.wav --state_dict S.pt [--filename out.json] [--chunk]

Same flags (`--audio/-a`, `--state_dict/-sd`, `--checkpoint/-c`, `--filename/-f`,
`--chunk_time`, `--step_time`, `--chunk`, `--plot`, all `--vap_*`), same JSON:
keys `probs, vad, p_now, p_future, H, loss` as nested lists. Files longer than
160 s switch to 25 s / 5 s chunked extraction like the reference (run.py:223-229).
`--precision fp32|bf16|fp16` is the one addition. CUDA is required.
"""
from __future__ import annotations

from argparse import ArgumentParser
from os.path import basename

import torch

from .audio import load_waveform
from .model import VapConfig, VapGPT
from .session import step_extraction
from .utils import batch_to_device, tensor_dict_to_json, write_json


def get_args(argv=None):
    parser = ArgumentParser()
    parser.add_argument("-a", "--audio", type=str, help="Path to waveform", required=True)
    parser.add_argument("-sd", "--state_dict", type=str,
                        default="example/VAP_3mmz3t0u_50Hz_ad20s_134-epoch9-val_2.56.pt",
                        help="Path to state_dict")
    parser.add_argument("-c", "--checkpoint", type=str, default=None, help="Path to trained model")
    parser.add_argument("-f", "--filename", type=str, default=None, help="Path to output json")
    parser.add_argument("--chunk_time", type=float, default=20, help="Duration of each chunk processed by model")
    parser.add_argument("--step_time", type=float, default=5, help="Increment to process in a step")
    parser.add_argument("--chunk", action="store_true", help="Process the audio in chunks (longer > 164s on 24Gb GPU audio)")
    parser.add_argument("--plot", action="store_true", help="Visualize output (matplotlib)")
    parser.add_argument("--precision", default=None, choices=["fp32", "bf16", "fp16"])
    parser, _ = VapConfig.add_argparse_args(parser, [])
    args = parser.parse_args(argv)
    return args, VapConfig.args_to_conf(args)


def infer(args, conf):
    """Everything run.py's main does up to (not including) writing the file."""
    if args.checkpoint is not None:
        raise NotImplementedError("Not implemeted from checkpoint...")  # run.py:206, verbatim behaviour
    if not torch.cuda.is_available():
        raise RuntimeError("voiceactivityprojection_b200 needs a CUDA device (no CPU fallback)")
    model = VapGPT(conf, precision=args.precision)
    model.load_state_dict(torch.load(args.state_dict, map_location="cpu"))
    model = model.to("cuda").eval()
    waveform, _ = load_waveform(args.audio, sample_rate=model.sample_rate)
    duration = round(waveform.shape[-1] / model.sample_rate)
    if waveform.shape[0] == 1:
        waveform = torch.cat((waveform, torch.zeros_like(waveform)))
    waveform = waveform.unsqueeze(0)
    if duration > 160:
        args.chunk = True
    if args.chunk:
        # like the reference, --chunk_time / --step_time are parsed but the defaults are used (run.py:236)
        out = step_extraction(waveform, model, "cuda")
    else:
        out = batch_to_device(model.probs(waveform.to("cuda")), "cpu")
    return out, waveform


def main(argv=None):
    args, conf = get_args(argv)
    out, _ = infer(args, conf)
    for k, v in out.items():
        if isinstance(v, torch.Tensor):
            print(f"{k}: ", tuple(v.shape))
    if args.filename is None:
        args.filename = basename(args.audio).replace(".wav", ".json")
    if not args.filename.endswith(".json"):
        args.filename += ".json"
    write_json(tensor_dict_to_json(out), args.filename)
    print("wavefile: ", args.audio)
    print("Saved output -> ", args.filename)
    if args.plot:
        raise NotImplementedError("--plot: plotting (vap/plot_utils.py) is outside the inference path")


if __name__ == "__main__":
    main()
