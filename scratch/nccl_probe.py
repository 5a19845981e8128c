import os, time, torch, torch.distributed as dist
rank=int(os.environ['RANK']); lr=int(os.environ['LOCAL_RANK']); torch.cuda.set_device(lr)
dev=torch.device('cuda',lr)
dist.init_process_group('nccl', device_id=dev)
x=torch.ones(10,dtype=torch.float64,device=dev)
for _ in range(5): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
ts=[]
for i in range(20):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    t0=time.perf_counter(); e0.record(); dist.all_reduce(x); e1.record(); torch.cuda.synchronize(); t1=time.perf_counter()
    ts.append((e0.elapsed_time(e1), (t1-t0)*1e3))
if rank==0: print('allreduce ms (event, wall):', [(round(a,3),round(b,3)) for a,b in ts[:8]])
# with some GPU work in between on the compute stream
a=torch.randn(4096,4096,device=dev)
ts=[]
for i in range(10):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); b=a@a; s=b.sum().double().reshape(1).repeat(10); dist.all_reduce(s); y=s+1; e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
if rank==0: print('matmul+allreduce ms:', [round(a,3) for a in ts])
dist.destroy_process_group()
