import torch, time
dev='cuda:0'
BV,Cin,HW,Cout=32,256,9216,32
x=torch.randn(BV,Cin,96,96,device=dev); conv=torch.nn.Conv2d(Cin,Cout,1).to(dev)
w=conv.weight.detach().view(Cout,Cin); b=conv.bias.detach()
def t(fn,name):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize(); print('%-50s %.1f us'%(name,e0.elapsed_time(e1)*100))
with torch.no_grad():
  for tf in (False,True):
    torch.backends.cuda.matmul.allow_tf32=tf; torch.backends.cudnn.allow_tf32=tf
    t(lambda: conv(x), 'cudnn conv NCHW tf32=%s'%tf)
    t(lambda: torch.baddbmm(b.view(1,1,Cout), x.view(BV,Cin,HW).transpose(1,2), w.t().unsqueeze(0).expand(BV,Cin,Cout)), 'baddbmm TT expand tf32=%s'%tf)
    t(lambda: torch.matmul(x.view(BV,Cin,HW).transpose(1,2), w.t()), 'matmul (view^T @ w^T) tf32=%s'%tf)
    wc=w.t().contiguous()
    t(lambda: torch.matmul(x.view(BV,Cin,HW).transpose(1,2), wc), 'matmul (view^T @ wc) tf32=%s'%tf)
    t(lambda: torch.matmul(w, x.view(BV,Cin,HW)), 'matmul NCHW out (w @ x) tf32=%s'%tf)
    t(lambda: torch.einsum('bcp,oc->bpo', x.view(BV,Cin,HW), w), 'einsum bcp,oc->bpo tf32=%s'%tf)
    xcl=x.contiguous(memory_format=torch.channels_last)
    t(lambda: conv(xcl), 'cudnn conv channels_last input tf32=%s'%tf)
    t(lambda: x.contiguous(memory_format=torch.channels_last), 'NCHW->NHWC copy of the 256-ch input')
