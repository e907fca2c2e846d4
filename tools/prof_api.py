import cProfile, pstats, numpy as np, torch, sys, io
sys.path.insert(0,'.')
import ssqueezepy_b200 as S
from oracle import ssq_oracle as O
N=160000
x=O.chirp(N,0); xd=torch.as_tensor(x,device='cuda')
w=S.Wavelet('morlet'); scales=O.bench_scales(O.OracleWavelet('morlet','float32'),N,300)
for _ in range(3): S.ssq_cwt(xd,w,scales=scales)
torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable()
for _ in range(20): S.ssq_cwt(xd,w,scales=scales)
torch.cuda.synchronize()
pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('cumulative').print_stats(22); print(s.getvalue()[:3500])
