"""Developer tool (GPU): copy bandwidth and bf16 matmul rate of the box, to tell a slow box from a slow kernel (compare with MEASURED_PEAKS.json)."""
import torch, time, subprocess
a=torch.empty(1<<30,dtype=torch.bfloat16,device='cuda'); b=torch.empty_like(a)
best=1e9
for i in range(10):
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record(); b.copy_(a); e.record(); torch.cuda.synchronize(); best=min(best,s.elapsed_time(e))
print("copy GB/s", 2*a.numel()*2/best/1e6)
x=torch.randn(8192,8192,dtype=torch.bfloat16,device='cuda'); y=torch.randn(8192,8192,dtype=torch.bfloat16,device='cuda')
best=1e9
for i in range(10):
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record(); z=x@y; e.record(); torch.cuda.synchronize(); best=min(best,s.elapsed_time(e))
print("bf16 TF/s", 2*8192**3/best/1e9)
# a latency-bound dependent chain on one thread per SM-ish: measures the effective SM clock
