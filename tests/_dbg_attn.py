import sys, torch
sys.path.insert(0,'/root/repo')
from lightglue_b200 import LightGlue, synth
from oracle import lightglue_oracle as oracle
torch.set_grad_enabled(False)
sd=synth.make_state_dict()
for (m_,n) in ((128,192),(192,128),(192,192)):
    data,_=synth.make_pair(n, m=m_, seed=1000)
    ref=oracle.forward(sd,data)
    m=LightGlue(features=None, depth_confidence=-1, width_confidence=-1, precision='bf16x3'); m.load_state_dict(sd,strict=False); m=m.cuda()
    out=m({k:{kk:vv.cuda() for kk,vv in v.items()} for k,v in data.items()})
    torch.cuda.synchronize()
    code=m.debug_timeout_code()
    print(m_, n, 'timeout code', hex(code), [ (i,hex(w)) for i,w in enumerate(m.debug_words) if w], 'flips', int((out['matches0'].cpu()!=ref['matches0']).sum()), flush=True)
