"""Host->device copy bandwidth of a pinned 617 MB buffer as a function of the NUMA placement of the host pages."""
import glob
import os
import time

import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out += list(range(int(a), int(b or a) + 1))
    return out


p = torch.cuda.get_device_properties(0)
bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
loc = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
print("gpu0", bus, "numa_node", node, "local_cpulist", loc, "cpu_count", os.cpu_count(), "affinity now", len(os.sched_getaffinity(0)))
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
dev = torch.device("cuda", 0)
dst = torch.empty(617 * 2 ** 20 // 4, device=dev)
all_cpus = sorted(os.sched_getaffinity(0))
for nd in nodes + ["default"]:
    if nd == "default":
        os.sched_setaffinity(0, all_cpus)
    else:
        cpus = [c for c in cpulist(open(nd + "/cpulist").read()) if c in all_cpus]
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
    src = torch.empty(617 * 2 ** 20 // 4)
    src.fill_(1.0)                      # first touch on the bound node
    src = src.pin_memory()
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(os.path.basename(nd), f"{src.numel() * 4 / dt / 1e9:.1f} GB/s", f"{dt * 1e3:.1f} ms", flush=True)
    del src
