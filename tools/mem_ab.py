import sys, torch, json
sys.path.insert(0, ".")
import ammcnet_aaai2021_b200 as A
from ammcnet_aaai2021_b200 import synth, functions as F_
dev="cuda:0"
C,D,M,k,b=512,64,256,2,64
p=synth.memory_params(3,C,D,M,k)
m=A.enc_quan_dec_res_topk(C,D,M,k=k); m.load_state_dict({"quan."+kk:v for kk,v in p.items()}); m=m.to(dev).eval()
x=synth.features(7,b,C,32,32).to(dev)
def timed(fn,n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a,b_=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b_.record(); torch.cuda.synchronize()
    return a.elapsed_time(b_)/n
with torch.no_grad():
    t=timed(lambda: m(x))
print(json.dumps({"module_ms": t}))
