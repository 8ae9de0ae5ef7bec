timeout 600 ncu --set full --clock-control none -k regex:pyr_stage -s 8 -c 2 -o gpurun_out/r02_pyr_f148 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-suite --no-sharded --batch-pairs 0 > gpurun_out/ncu_pyr.log 2>&1
tail -2 gpurun_out/ncu_pyr.log
python - <<P
import torch, time
x=torch.empty(1<<30,dtype=torch.uint8,device='cuda')
for name,fn in (('memset 1GiB',lambda: x.zero_()),('fill f32',lambda: x.view(torch.float32).fill_(1.5))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record(); 
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/10
    print(name,'%.3f ms %.1f GB/s'%(ms,(1<<30)/ms/1e6))
y=torch.empty(1<<30,dtype=torch.uint8,device='cuda')
for _ in range(3): y.copy_(x)
torch.cuda.synchronize(); a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True); a.record()
for _ in range(10): y.copy_(x)
b.record(); torch.cuda.synchronize(); ms=a.elapsed_time(b)/10; print('copy 1GiB %.3f ms %.1f GB/s (r+w)'%(ms,2*(1<<30)/ms/1e6))
P
