import torch
x=torch.empty(1<<31,dtype=torch.uint8,device='cuda'); y=torch.empty_like(x)
def t(fn,n=5):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
ms=t(lambda: x.fill_(1)); print("fill 2GiB: %.3f ms  %.0f GB/s write"%(ms, x.numel()/ms/1e6))
ms=t(lambda: y.copy_(x)); print("copy 2GiB: %.3f ms  %.0f GB/s r+w"%(ms, 2*x.numel()/ms/1e6))
xs=x[:1<<29]
ms=t(lambda: xs.fill_(1)); print("fill 512MiB: %.3f ms  %.0f GB/s write"%(ms, xs.numel()/ms/1e6))
