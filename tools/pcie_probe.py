"""PCIe copy bandwidth between pinned host memory and the GPU (cudaMemcpyAsync through torch), warm, several sizes."""
import torch, time
for mb in (2, 8, 32, 128):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        d.copy_(h, non_blocking=True); h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    R = 40
    e[0].record()
    for _ in range(R): d.copy_(h, non_blocking=True)
    e[1].record()
    for _ in range(R): h.copy_(d, non_blocking=True)
    e[2].record(); torch.cuda.synchronize()
    print(f"{mb:4d} MB: H2D {R * n / e[0].elapsed_time(e[1]) / 1e6:6.1f} GB/s   D2H {R * n / e[1].elapsed_time(e[2]) / 1e6:6.1f} GB/s")
