import sys, os, numpy as np; sys.path.insert(0,'.')
import torch, uavenv_b200 as ub
fx=np.load("tests/golden/policy_net.npz")
net=ub.TransformerActorCritic().cuda().eval()
net.load_state_dict({str(k): torch.from_numpy(fx["p::"+str(k)]) for k in fx["keys"]})
for B in (25, 96, 1000):
    obs=torch.rand(B,5,14,device="cuda"); obs[:,:,13]=1
    for b in range(B): obs[b,:b%5]=0
    f=ub.FusedPolicyForward(B,"cuda"); f.sync(net)
    a,lp,v,e=f.get_action(obs,1); torch.cuda.synchronize()
    with torch.no_grad(): rl,rv=net.logits_and_value(obs)
    print(B,"logit err",float((f.logits[:B]-rl).abs().max()),"value err",float((v-rv).abs().max()), "ref v max", float(rv.abs().max()))
    f.close()
