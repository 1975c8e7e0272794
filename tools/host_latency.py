"""Host <-> GPU round-trip latencies of this box (launch + sync, small D2H), to tell a slow host from a slow kernel."""
import time
import torch
x = torch.zeros(1, device="cuda")
h = torch.zeros(1, pin_memory=True)
torch.cuda.synchronize()
for name, fn in (("launch+sync", lambda: (x.add_(1), torch.cuda.synchronize())),
                 ("launch+4B D2H (pinned)+sync", lambda: (x.add_(1), h.copy_(x, non_blocking=True), torch.cuda.synchronize())),
                 ("launch+.item()", lambda: x.add_(1).item())):
    for _ in range(50):
        fn()
    t = time.perf_counter()
    for _ in range(500):
        fn()
    print("%-32s %.1f us" % (name, (time.perf_counter() - t) / 500 * 1e6))
import os
print("cpus", os.cpu_count(), "load", os.getloadavg())
