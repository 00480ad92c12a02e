import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from octcubem_b200 import ops
from octcubem_b200._lib import OCT_BF16
B, S, H, d = 8, 4097, 16, 32
dev = torch.device("cuda:0")
qkv = (torch.randn(B, S, 3 * H * d, device=dev) * 0.5).bfloat16()
dout = torch.randn(B, S, H * d, device=dev).bfloat16()
out, lse = ops.attn_fwd(qkv, H, d, OCT_BF16)
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
bw = lambda: ops.attn_bwd(qkv, out, dout, lse, H, d, OCT_BF16)
fw = lambda: ops.attn_fwd(qkv, H, d, OCT_BF16)
print("cold: fwd %.1f bwd %.1f" % (timeit(fw), timeit(bw)))
a = torch.randn(8192, 8192, device=dev).bfloat16(); b = torch.randn(8192, 8192, device=dev).bfloat16()
t0 = time.time()
while time.time() - t0 < 4.0:
    for _ in range(20): a @ b
    torch.cuda.synchronize()
print("after 4 s of GEMM burn: fwd %.1f bwd %.1f" % (timeit(fw), timeit(bw)))
print("n=100: bwd %.1f fwd %.1f" % (timeit(bw, 100), timeit(fw, 100)))
print("n=400: bwd %.1f" % (timeit(bw, 400)))
os.system("nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader")
