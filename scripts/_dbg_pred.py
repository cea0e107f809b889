import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from pbnn_b200.ops import FlatSpec, default_ops
ops = default_ops()
if len(sys.argv) > 3: ops.set_option('debug', int(sys.argv[3]))
S, M = int(sys.argv[1]), int(sys.argv[2])
layers = (9, 100, 100, 1)
fs = FlatSpec(layers, 0, 0.1)
torch.manual_seed(0)
smp = (0.1 * torch.randn(S, fs.num_params, device="cuda")).contiguous()
Xt = torch.randn(M, layers[0], device="cuda")
m = ops.predict(fs, smp, Xt, want_preds=True, want_moments=True); torch.cuda.synchronize()
print(S, M, "ok", float(m["mean"].abs().max()))
