import torch, time
x = torch.empty(512 * 1024 * 1024 // 4).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(2):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    d.copy_(x, non_blocking=True)
e1.record()
torch.cuda.synchronize()
print("H2D pinned GB/s", 5 * x.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
t = time.perf_counter(); y = x.to("cuda", non_blocking=True); torch.cuda.synchronize(); print("to() ms", (time.perf_counter() - t) * 1e3)
