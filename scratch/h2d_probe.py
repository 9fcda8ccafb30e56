import torch, time
n=4096
h=torch.empty((n,5,256,256),dtype=torch.float32).pin_memory()
d=torch.empty_like(h,device='cuda')
for chunk in (4096,512,128):
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e0.record()
        for a in range(0,n,chunk): d[a:a+chunk].copy_(h[a:a+chunk],non_blocking=True)
        e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1); print('chunk',chunk,'H2D GB/s',h.numel()*4/ms/1e6)
