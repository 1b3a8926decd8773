import torch, time
dev='cuda'
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
out=torch.empty(4096*3*224*224, dtype=torch.float32, device=dev)
src=torch.empty_like(out)
ms=t(lambda: out.fill_(1.5)); print("fill 2.47GB: %.3f ms  %.0f GB/s"%(ms, out.numel()*4/ms/1e6))
ms=t(lambda: out.copy_(src)); print("copy 2.47GB->2.47GB: %.3f ms  %.0f GB/s (r+w)"%(ms, 2*out.numel()*4/ms/1e6))
u8=torch.empty(64*1080*1920*3, dtype=torch.uint8, device=dev)
ms=t(lambda: u8.sum()); print("read 398MB u8 sum: %.3f ms %.0f GB/s"%(ms, u8.numel()/ms/1e6))
h=out[:out.numel()//2]
ms=t(lambda: h.fill_(2.0)); print("fill 1.23GB: %.3f ms  %.0f GB/s"%(ms, h.numel()*4/ms/1e6))
