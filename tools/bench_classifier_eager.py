import torch, torch.nn.functional as F, json
def run(n,in_f,k,train,dtype=torch.float32):
    w=torch.randn(512,in_f,device='cuda',dtype=dtype,requires_grad=train); b=torch.randn(512,device='cuda',dtype=dtype,requires_grad=train)
    E=(F.normalize(torch.randn(k-1,512,device='cuda'))*0.8).to(dtype); bg=torch.randn(1,512,device='cuda',dtype=dtype,requires_grad=train)
    x=torch.randn(n,in_f,device='cuda',dtype=dtype,requires_grad=train); labels=torch.randint(0,k-1,(n,),device='cuda')
    def step():
        h=F.normalize(F.linear(x,w,b))
        emb=torch.cat([E,F.normalize(bg)])
        y=h@emb.T
        y=y*50.0-3.0
        if train: F.cross_entropy(y.float(),labels).backward()
    for _ in range(5): step()
    torch.cuda.synchronize()
    a,bb=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): step()
    bb.record(); torch.cuda.synchronize()
    return dict(n=n,in_f=in_f,k=k,train=train,dtype=str(dtype),us=a.elapsed_time(bb)/50*1e3)
for dt in (torch.float32, torch.float16):
    for args in ((1000,1024,66,False),(1000,1024,1204,False),(1678,1024,1204,True),(2,256,65,False)):
        print(json.dumps(run(*args,dtype=dt)))
