import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'..')
import numpy as np, torch
from oracle import pbnn_oracle as o
from pbnn_b200.ops import FlatSpec, default_ops
ops = default_ops()
def timeit(f, n=3):
    f(); torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
# cfg2 HMC
layers=(13,50,1); N=506; C=256; L=20; S=10
X,y=o.synthetic_regression(N,13,0); fs=FlatSpec(layers,0,0.1); sp=o.MlpSpec(layers,0,0.1)
keys=o.split(o.PRNGKey(0),C); th0=o.init_positions(sp,keys)
Xd,yd,thd,kd = (ops._as(a, torch.float32) for a in (X,y,th0)) .__iter__().__next__() if False else (None,None,None,None)
Xd=ops._as(X,torch.float32); yd=ops._as(y,torch.float32); thd=ops._as(th0,torch.float32); kd=ops._keys(keys)
minv=torch.ones(sp.num_params,device='cuda')
ms=timeit(lambda: ops.hmc(fs,Xd,yd,thd,kd,S,1e-3,minv,L,keep_trace=False))
fl=C*S*L*1467400
print(f"cfg2 HMC: {ms:.3f} ms for {C}x{S} steps -> {C*S/ms*1e3:.0f} chain-steps/s, {fl/ms/1e9:.2f} TFLOP/s", ops.plan(fs,N,6))
# cfg1 SGLD single chain
layers=(8,50,1); N=1030; T=2000
X,y=o.synthetic_regression(N,8,0); fs=FlatSpec(layers,0,0.1); sp=o.MlpSpec(layers,0,0.1)
keys=o.split(o.PRNGKey(0),1); th0=o.init_positions(sp,keys)
Xd=ops._as(X,torch.float32); yd=ops._as(y,torch.float32)
ms=timeit(lambda: ops.sgld(fs,Xd,yd,th0,keys,32,1e-5,T,keep_trace=False))
print(f"cfg1 SGLD 1 chain: {ms/T*1e3:.2f} us/step -> {T/ms*1e3:.0f} steps/s")
for C in (148, 1024):
    keys=o.split(o.PRNGKey(0),C); th0=o.init_positions(sp,keys)
    ms=timeit(lambda: ops.sgld(fs,Xd,yd,th0,keys,32,1e-5,T,keep_trace=False))
    print(f"cfg1-shape SGLD {C} chains: {C*T/ms*1e3:.0f} steps/s, {C*T*60800/ms/1e9:.2f} TFLOP/s")
# cfg3 pSGLD / SGHMC
layers=(9,100,100,1); N=45730; C=1024; T=50
X,y=o.synthetic_regression(N,9,0); fs=FlatSpec(layers,0,0.1); sp=o.MlpSpec(layers,0,0.1)
keys=o.split(o.PRNGKey(0),C); th0=o.init_positions(sp,keys)
Xd=ops._as(X,torch.float32); yd=ops._as(y,torch.float32)
for nm,f,fl in (("sgld", lambda: ops.sgld(fs,Xd,yd,th0,keys,128,1e-6,T,keep_trace=False), 8.218e6),
             ("psgld", lambda: ops.psgld(fs,Xd,yd,th0,keys,128,1e-6,T,0.95,keep_trace=False), 8.218e6),
             ("sghmc", lambda: ops.sghmc(fs,Xd,yd,th0,keys,128,1e-9,T//5,10,keep_trace=False), 82.18e6/5)):
    ms=timeit(f,2)
    print(f"cfg3 {nm}: {C*T/ms*1e3:.0f} (sgld-equiv) steps/s, {C*T*fl/ms/1e9:.2f} TFLOP/s", ops.plan(fs,128,3))
# cfg4 adam
layers=(8,50,50,1); N=9568; C=512
X,y=o.synthetic_regression(N,8,0); fs=FlatSpec(layers,0,0.1); sp=o.MlpSpec(layers,0,0.1)
keys=o.split(o.PRNGKey(0),C); th0=o.init_positions(sp,keys)
idx=torch.randint(0,N,(C,299,32),dtype=torch.int32,device='cuda')
Xd=ops._as(X,torch.float32); yd=ops._as(y,torch.float32)
ms=timeit(lambda: ops.adam_ensemble(fs,Xd,yd,th0,idx,1e-2),2)
print(f"cfg4 adam: {C*299/ms*1e3:.0f} steps/s, {C*299*540800/ms/1e9:.2f} TFLOP/s")
ms=timeit(lambda: ops.permutation(o.split(o.PRNGKey(1),512*11), N),2)
print(f"cfg4 permutations 512x11 of {N}: {ms:.2f} ms")
samples=torch.randn(512, sp.num_params, device='cuda')*0.1; Xt=torch.randn(1024,8,device='cuda')
ms=timeit(lambda: ops.predict(fs,samples,Xt,want_preds=False,want_moments=True),3)
print(f"cfg4 predict 512x1024: {ms:.3f} ms, {512*1024*2*2950/ms/1e9:.2f} TFLOP/s")
