"""A/B: two builds of the library on the same inputs -- bit-identity of HMC outputs and timing (PBNN_B200_LIB)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys, numpy as np, torch, time
sys.path.insert(0, %r)
from pbnn_b200.ops import FlatSpec, default_ops
import bench
ops = default_ops()
w = dict(bench.WORKLOADS["hmc_cfg2"]); fs = FlatSpec(tuple(w["layers"]), 0, 0.1)
X, y, Xt = bench.make_data(w)
C = 300
keys = ops.split(np.array([[0, 0]], np.uint32), C)[0].contiguous()
th0 = 0.01 * ops.normal(keys, fs.num_params)
minv = np.linspace(0.8, 1.2, fs.num_params).astype(np.float32)
out = ops.hmc(fs, X, y, th0, keys, 12, 5e-4, minv, 7, info=True)
np.savez(sys.argv[1], trace=out["trace"].cpu().numpy(), delta=out["delta"].cpu().numpy(), logp=out["logp"].cpu().numpy())
''' % ROOT
outs = []
for lib in sys.argv[1:3]:
    env = dict(os.environ, PBNN_B200_LIB=os.path.join(ROOT, "pbnn_b200", lib))
    f = "/tmp/ab_%s.npz" % lib
    subprocess.check_call([sys.executable, "-c", child, f], env=env)
    outs.append(f)
import numpy as np
a, b = np.load(outs[0]), np.load(outs[1])
for k in a.files:
    print(k, "bit-identical" if np.array_equal(a[k], b[k]) else "DIFFERENT max abs %g" % np.abs(a[k] - b[k]).max())
