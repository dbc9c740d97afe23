"""Measured dense TF32 tensor-core peak of this GPU (the KSD / K2 roofline denominator; MEASURED_PEAKS.json has none):
cuBLAS TF32 GEMM (torch.matmul, fp32 inputs, allow_tf32) 8192^3, best of 10 (burst) and back to back for 3 s (sustained),
CUDA events - the same recipe MEASURED_PEAKS.json states for bf16.  Writes profiles/tf32_peak.json."""
import json, os, sys, time
import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2.0 * n ** 3 / (best * 1e-3) / 1e12
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps, t0 = 0, time.perf_counter()
e0.record()
while time.perf_counter() - t0 < 3.0:
    for _ in range(20):
        c = a @ b
    reps += 20
e1.record(); torch.cuda.synchronize()
sustained = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
# bf16 by the same recipe, for the ratio on THIS box
ab, bb = a.bfloat16(), b.bfloat16()
for _ in range(3):
    cb = ab @ bb
torch.cuda.synchronize()
bestb = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cb = ab @ bb; e1.record(); torch.cuda.synchronize()
    bestb = min(bestb, e0.elapsed_time(e1))
out = {"tf32_tflops": burst, "tf32_tflops_sustained": sustained, "bf16_tflops_same_box": 2.0 * n ** 3 / (bestb * 1e-3) / 1e12,
       "how": "torch.matmul fp32 with allow_tf32 (cuBLAS TF32) 8192^3: best of 10 (burst), back to back for 3 s (sustained), CUDA events",
       "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
print(json.dumps(out))
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "tf32_peak.json")
os.makedirs(os.path.dirname(dst), exist_ok=True)
json.dump(out, open(dst, "w"), indent=1)
