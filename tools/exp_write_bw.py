import torch, time
x = torch.empty(1 << 30, dtype=torch.int32, device="cuda")  # 4 GiB
y = torch.empty(1 << 30, dtype=torch.int32, device="cuda")
for name, fn, nbytes in (("fill (write only)", lambda: x.fill_(7), x.numel()*4), ("copy (read+write)", lambda: y.copy_(x), 2*x.numel()*4), ("sum (read only)", lambda: x.sum(), x.numel()*4)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)/10
    print(f"{name}: {nbytes/ms/1e6:.0f} GB/s")
