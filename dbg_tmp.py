import sys; sys.path.insert(0,'.')
import torch, uavenv_b200 as ub
from torch.profiler import profile, ProfilerActivity
net=ub.TransformerActorCritic().cuda()
n=131072
obs=torch.rand(n,5,14,device="cuda"); act=torch.randint(0,2,(n,),device="cuda")
torch.backends.cuda.matmul.allow_tf32=True
def step():
    lp,v,e=net.evaluate(obs,act); loss=(lp.mean()+v.mean()+e.mean()); loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
