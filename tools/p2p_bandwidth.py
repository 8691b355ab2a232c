"""Raw NVLink peer bandwidth of the box (torch copies between two GPUs): the yardstick for the tensor-parallel
exchange numbers in DESIGN.md section 4 (measured: 734 GB/s for cudaMemcpyPeer 1 -> 0 on 2 x B200)."""
import torch, time
n = torch.cuda.device_count()
print("gpus", n)
for i in range(n):
    for j in range(n):
        if i != j:
            print(i, j, torch.cuda.can_device_access_peer(i, j), end=" | ")
print()
sz = 256 << 20
a = [torch.empty(sz, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
b = [torch.empty(sz, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
def sync():
    for i in range(n): torch.cuda.synchronize(i)
# one direction 1 -> 0 (copy engine)
for rep in range(2):
    sync(); t=time.perf_counter()
    for _ in range(10): b[0].copy_(a[1], non_blocking=True)
    sync(); dt=time.perf_counter()-t
    print("memcpy 1->0", sz*10/dt/1e9, "GB/s")
# all-to-next ring simultaneously
sync(); t=time.perf_counter()
for _ in range(10):
    for i in range(n):
        with torch.cuda.device(i):
            b[i].copy_(a[(i+1)%n], non_blocking=True)
sync(); dt=time.perf_counter()-t
print("ring pull per gpu", sz*10/dt/1e9, "GB/s")
