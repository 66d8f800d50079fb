import torch,time
for mb in (64,256,1024):
    a=torch.empty(mb*1024*1024,dtype=torch.uint8,pin_memory=True); d=torch.empty_like(a,device='cuda')
    for name,src,dst in (("H2D",a,d),("D2H",d,a)):
        dst.copy_(src,non_blocking=True); torch.cuda.synchronize()
        t=time.perf_counter()
        for _ in range(3): dst.copy_(src,non_blocking=True)
        torch.cuda.synchronize(); el=(time.perf_counter()-t)/3
        print(mb,"MB",name,f"{mb/1024/el:.1f} GB/s")
