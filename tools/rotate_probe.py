import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from toucan_b200 import capi, ops
capi.check(capi.lib().tc_set_device(0), "dev")
a = torch.rand((2160, 3840, 4), device="cuda")
for i in range(4):
    out = ops.rotate(a, 45.0)
torch.cuda.synchronize()
print("ok", float(out[1000, 1000, 0]))
