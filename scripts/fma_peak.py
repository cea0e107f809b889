import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from pbnn_b200.ops import default_ops
ops = default_ops(); lib = ops.lib
lib.pbnn_probe_fp32.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
sink = torch.zeros(4, device="cuda")
for mode in (0, 1):
    for ctas in (148 * 4, 148 * 8):
        iters = 20000
        lib.pbnn_probe_fp32(mode, 100, ctas, sink.data_ptr(), None); torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); lib.pbnn_probe_fp32(mode, iters, ctas, sink.data_ptr(), None); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"mode {'FFMA2' if mode else 'FFMA'} ctas {ctas}: {ctas*256*iters*64/ms/1e9:.2f} TFLOP/s ({ms:.2f} ms)")
