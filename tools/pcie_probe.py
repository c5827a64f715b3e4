"""PCIe D2H probe: pinned host buffer, 1/2/4 concurrent streams (what bounds bench.py's e2e)."""
import time, torch
n = 3 * 1024**3  # bytes
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
for streams in (1, 2, 4, 8):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    chunk = n // streams
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for i, s in enumerate(ss):
            with torch.cuda.stream(s):
                h[i * chunk:(i + 1) * chunk].copy_(d[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    print(f"D2H streams={streams}: {n / dt / 1e9:.1f} GB/s")
t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); print(f"H2D: {n/(time.perf_counter()-t)/1e9:.1f} GB/s")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
print(subprocess.run(["bash", "-c", "lscpu | egrep 'Model name|Socket|NUMA node\\(s\\)|^CPU\\(s\\)'; free -g | head -2"], capture_output=True, text=True).stdout)
