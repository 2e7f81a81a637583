import torch,time
x=torch.empty(2*1024**3,dtype=torch.int32,device='cuda').fill_(1)  # 8 GB
for f,name in [(lambda: x.sum(), 'sum int32'), (lambda: x.view(torch.float32).max(), 'max f32'), (lambda: torch.count_nonzero(x), 'count_nonzero')]:
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(name, 'GB/s', x.numel()*4/ms/1e6)
y=torch.empty_like(x)
for _ in range(2): y.copy_(x)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y.copy_(x)
e1.record(); torch.cuda.synchronize()
print('copy GB/s (r+w)', 2*x.numel()*4/(e0.elapsed_time(e1)/5)/1e6)
