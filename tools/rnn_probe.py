"""Times the tensor-core recurrence alone and dumps SM-clock samples of its phases
(cluster 0, rank 0, group 0, steps 64..95)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from voiceactivityprojection_b200 import _lib

lib = _lib.load()
nseq, T = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 2000
GROUPS = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # shape: 100*NB+NG
torch.manual_seed(0)
cell = torch.nn.LSTM(256, 256, batch_first=True)
wcat = torch.empty((1024, 512)); bias = torch.empty(1024)
lib.vapb_debug_rnn_pack(0, cell.weight_ih_l0.detach().contiguous().data_ptr(), cell.weight_hh_l0.detach().contiguous().data_ptr(),
                        cell.bias_ih_l0.detach().contiguous().data_ptr(), cell.bias_hh_l0.detach().contiguous().data_ptr(), wcat.data_ptr(), bias.data_ptr())
wd, bd = wcat.cuda().bfloat16().contiguous(), bias.cuda()
x = (torch.randn(nseq, T, 256, device="cuda") * 0.7).bfloat16()
out = torch.zeros_like(x)
dbg = torch.zeros((32, 8), dtype=torch.int64, device="cuda")
err = C.create_string_buffer(512)
st = torch.cuda.current_stream().cuda_stream
def run(d):
    rc = lib.vapb_debug_rnn_tc(st, 0, x.data_ptr(), T * 256, 256, wd.data_ptr(), bd.data_ptr(), out.data_ptr(), T * 256, nseq, T, err, 512, d, GROUPS)
    assert rc == 0, err.value
for _ in range(2): run(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(None); run(None); e1.record(); torch.cuda.synchronize()
print("ms per launch", e0.elapsed_time(e1) / 2, "us/step", e0.elapsed_time(e1) / 2 / T * 1e3)
run(dbg.data_ptr()); torch.cuda.synchronize()
d = dbg.cpu()
base = d[:, 0:1]
names = ["mma:hfull", "mma:issued", "gate:accfull", "gate:act done", "gate:bar1", "gate:bar2", "gate:copies issued"]
print(names)
for i in range(8, 16):
    print(i + 64, [int(v) for v in (d[i, :7] - d[i, 0])], "next hfull", int(d[i + 1, 0] - d[i, 0]))
