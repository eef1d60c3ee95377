import sys, time, torch, collections
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tests')
import __graft_entry__ as g
g.build(); pkg = g.load_package()
from oracle import unet as O
from golden_configs import GOLDEN_CONFIGS
cfg = GOLDEN_CONFIGS["cifar"][0]
params = O.seeded_params(cfg, 0)
m = pkg.UNetModelWrapper(dim=(3,32,32), num_res_blocks=2, num_channels=128, channel_mult=[1,2,2,2], num_heads=4, num_head_channels=64, attention_resolutions="16", precision="bf16")
m.load_state_dict(params); m = m.cuda().eval()
B = int(sys.argv[1]) if len(sys.argv)>1 else 1024
x = torch.randn(B,3,32,32,device='cuda')
e = m.engine()
for _ in range(2): e.forward(x, 0.5)
torch.cuda.synchronize()
t0=time.time()
for _ in range(5): e.forward(x, 0.5)
torch.cuda.synchronize()
dt=(time.time()-t0)/5
print(f"B={B} NFE {dt*1e3:.2f} ms  -> {e.flops_per_sample*B/dt/1e12:.1f} TFLOP/s, ws={e.workspace_bytes(B)/2**30:.2f} GiB, launches={e.last_launches}")
import subprocess
def clk():
    return subprocess.run(["nvidia-smi","--query-gpu=clocks.sm,power.draw","--format=csv,noheader"],capture_output=True,text=True).stdout.strip()
print("clocks before profile:", clk())
rows = e.profile_forward(x, 0.5, repeats=10)
print("clocks after profile:", clk())
agg = collections.defaultdict(lambda:[0.0,0.0,0])
for r in rows:
    a=agg[r['kind']]; a[0]+=r['ms']; a[1]+=r['flops']; a[2]+=1
for k,(ms,fl,n) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:14s} n={n:3d} {ms:8.3f} ms  {fl/ms/1e9 if ms else 0:8.1f} TFLOP/s")
print("--- slowest ops")
for r in sorted(rows, key=lambda r:-r['ms'])[:25]:
    print(f"{r['name']:34s} {r['kind']:13s} {r['ms']:7.3f} ms {r['flops']/r['ms']/1e9 if r['ms'] else 0:8.1f} TF/s")

import os
os.makedirs('gpurun_out', exist_ok=True)
with open('gpurun_out/' + (sys.argv[2] if len(sys.argv) > 2 else 'profile_ops.txt'),'w') as f:
    for r in rows:
        f.write(f"{r['name']:36s} {r['kind']:13s} {r['ms']:8.4f} ms {r['flops']/r['ms']/1e9 if r['ms'] else 0:8.1f} TF/s\n")
