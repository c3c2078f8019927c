"""cuBLAS DGEMM yardstick (library call, measurement only): the FP64 denominator MEASURED_PEAKS.json lacks."""
import torch, time, json, subprocess
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_burst_tflops"] = 2 * n**3 / best * 1e-9
    # sustained ~4 s
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = max(3, int(4000 / best))
    e0.record()
    for _ in range(reps):
        c = a @ b
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res[f"dgemm_{n}_sustained_tflops"] = 2 * n**3 * reps / ms * 1e-9
    q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    res[f"smi_after_{n}"] = q
# syrk-like: A @ A.T
a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
c = a @ a.T; torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); c = a @ a.T; e1.record(); torch.cuda.synchronize()
res["dgemm_8192_NT_tflops"] = 2 * 8192**3 / e0.elapsed_time(e1) * 1e-9
print(json.dumps(res, indent=1))
