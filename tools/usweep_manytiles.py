import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from tensorbnn_b200 import _lib
from tensorbnn_b200.engine import Engine
import test_gpu_parity as tp
N=262181
arch, lik, X, Y, TH, HY = tp.problem("c2s", N)
res={}
for name, flags in (("umma", 8), ("ffma2", 0)):
    eng = Engine(arch, lik, dtype=torch.float32, chains=1, flags=flags); eng.set_data(X, Y)
    print(eng.sweep_info())
    lp,g,st = eng.logp_grad(TH,HY); res[name]=(lp.item(), g.cpu().numpy()[0], st.cpu().numpy())
    if name=="umma":
        h=N//2; parts=[]
        for sl in (slice(0,h), slice(h,N)):
            eng.set_data(X[sl],Y[sl]); parts.append(eng.logp_grad(TH,HY)[2].cpu().numpy())
        print("linearity", parts[0]+parts[1], res["umma"][2])
print("logp", res["umma"][0], res["ffma2"][0], abs(res["umma"][0]-res["ffma2"][0])/abs(res["ffma2"][0]))
gu,gf=res["umma"][1],res["ffma2"][1]
err=np.abs(gu-gf)/np.abs(gf).max(); i=err.argmax(); print("grad rel", err.max(), i, gu[i], gf[i])
r64=lambda a: np.asarray(a).astype(np.float32).astype(np.float64)
from oracle import analytic
lp_ref,g_ref=analytic.main_value_and_grad(arch, lik, r64(TH[0]), r64(HY[0]), r64(X), r64(Y))
for n in res: print(n, "vs oracle: logp", abs(res[n][0]-lp_ref)/abs(lp_ref), "grad", np.abs(res[n][1]-g_ref).max()/np.abs(g_ref).max())
