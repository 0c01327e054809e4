import subprocess, sys, os
CASES = [
 "x=torch.randn(4096,624,device='cuda'); w=torch.randn(400,624,device='cuda'); print(LA.gemm(x,w,trans_b=True).shape)",
 "x=torch.randn(4096,624,device='cuda',requires_grad=True); w=torch.randn(400,624,device='cuda',requires_grad=True); y=LA.linear(x,w,None); print(y.shape)",
 "x=torch.randn(4096,624,device='cuda',requires_grad=True); w=torch.randn(400,624,device='cuda',requires_grad=True); y=LA.linear(x,w,None); print(torch.autograd.grad(y,[x],torch.randn_like(y))[0].shape)",
 "x=torch.randn(4096,624,device='cuda',requires_grad=True); w=torch.randn(400,624,device='cuda',requires_grad=True); y=LA.linear(x,w,None); print(torch.autograd.grad(y,[w],torch.randn_like(y))[0].shape)",
]
for i,c in enumerate(CASES):
    code = "import sys; sys.path.insert(0,'.'); import torch; import recsys_benchmark_b200.linalg as LA; " + c + "; torch.cuda.synchronize(); print('OK')"
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    print("CASE", i, "rc", r.returncode, "|", r.stdout.strip()[-200:], "|", r.stderr.strip()[-3000:])
