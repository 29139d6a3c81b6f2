"""Developer scratch: HBM bandwidth of the box by read/write mix (library kernels: fill, copy, sum) - the
ceilings a write-heavy kernel (fk_jac: 28 B in, 196 B out) or a read-heavy one (reward: 60 in, 4 out) can reach."""
import torch
dev = torch.device("cuda")
n = 1 << 30  # 4 GiB of float32
a = torch.empty(n, dtype=torch.float32, device=dev); b = torch.empty(n, dtype=torch.float32, device=dev)
a.fill_(1.0); b.fill_(2.0)
def t(fn, rep=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(rep):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts) * 1e-3
gb = n * 4 / 1e9
print(f"write only (fill_)        : {gb / t(lambda: a.fill_(3.0)):8.0f} GB/s")
print(f"write only (cudaMemset)   : {gb / t(lambda: a.zero_()):8.0f} GB/s")
print(f"copy (read + write)       : {2 * gb / t(lambda: b.copy_(a)):8.0f} GB/s")
print(f"read only (sum)           : {gb / t(lambda: a.sum()):8.0f} GB/s")
c = torch.empty(n // 8, dtype=torch.float32, device=dev)
print(f"read 8 : write 1 (a[::8] strided gather excluded) -> add of two reads, one write: {3 * gb / t(lambda: torch.add(a, b, out=a)):8.0f} GB/s")
