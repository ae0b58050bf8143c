"""H2D bandwidth of separately allocated pinned host buffers (is the host-memory placement of this VM uniform?)."""
import torch, time
dev = torch.device("cuda", 0)
dst = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
bufs = []
for i in range(16):
    b = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
    bufs.append(b)
    dst.copy_(b, non_blocking=True); torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        dst.copy_(b, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    print(i, "GB/s %.1f" % (3 * 0.536870912 / (a.elapsed_time(e) / 1e3)), hex(b.data_ptr()))
