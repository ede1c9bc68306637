"""Measures dense GEMM throughput of cuBLAS on this GPU: tf32 (fp32 inputs, allow_tf32), bf16 and fp16, burst (best of 10)
and sustained (back to back for ~2 s).  Writes one JSON line; bench.py's executed-tensor fractions use the tf32 figure."""
import json, sys, time
import torch

def run(dtype, tf32, N=8192):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn((N, N), device="cuda", dtype=dtype)
    b = torch.randn((N, N), device="cuda", dtype=dtype)
    c = torch.empty((N, N), device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(10, int(2000.0 / best))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); torch.cuda.synchronize()
    fl = 2.0 * N ** 3
    return fl / (best * 1e-3) / 1e12, fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12

out = {}
for name, dt, tf in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, False), ("fp16", torch.float16, False)):
    b, s = run(dt, tf)
    out[name + "_tflops_burst"], out[name + "_tflops_sustained"] = round(b, 1), round(s, 1)
out["gpu"] = torch.cuda.get_device_name(0)
print(json.dumps(out))
