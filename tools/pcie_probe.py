import torch, time
n = 1 << 28  # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True); h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_a = torch.empty(n, dtype=torch.float32, device="cuda"); d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
gb = n * 4 / 1e9
print("H2D  GB/s", gb / t(lambda: d_a.copy_(h_in, non_blocking=True)))
print("D2H  GB/s", gb / t(lambda: h_out.copy_(d_b, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
print("duplex GB/s per direction", gb / t(both))
import subprocess; print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv,noheader", shell=True, capture_output=True, text=True).stdout)
