#!/usr/bin/env python
"""How much device-to-host bandwidth does the box give N GPUs at once?  (the end-to-end leg moves 117 GB of frames per GPU and
pass.)  usage: d2h_probe.py N [numa]   -- N processes, one per GPU, each copying 256 MB chunks into page-locked memory for 3 s;
with `numa`, a process first binds its memory to the NUMA node of its GPU (set_mempolicy), where sysfs names one."""
import os, sys, time, subprocess, ctypes

def gpu_numa_node(i):
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(i)], capture_output=True, text=True).stdout.strip().lower()
        bus = bus[4:] if len(bus) > 12 else bus          # 00000000:1B:00.0 -> 0000:1b:00.0
        return int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except Exception as e:
        return -1

def child(i, n, numa):
    node = gpu_numa_node(i)
    if numa and node >= 0:
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        r = libc.syscall(238, 2, ctypes.byref(mask), 64)   # set_mempolicy(MPOL_BIND, mask, maxnode)
        if r != 0: print(f"gpu {i}: set_mempolicy failed ({ctypes.get_errno()})", flush=True)
    import torch
    torch.cuda.set_device(i)
    src = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dst = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    dst.copy_(src); torch.cuda.synchronize()
    # all processes start together
    while time.time() < float(os.environ["D2H_START"]): pass
    t0 = time.time(); nb = 0
    while time.time() - t0 < 3.0:
        for _ in range(4): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize(); nb += 4 * (256 << 20)
    dt = time.time() - t0
    print(f"gpu {i} (numa node {node}, bind {int(numa)}): {nb / dt / 1e9:.1f} GB/s", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "child":
        child(int(sys.argv[1]), 0, sys.argv[2] == "1")
    else:
        n = int(sys.argv[1]); numa = len(sys.argv) > 2 and sys.argv[2] == "numa"
        os.environ["D2H_START"] = str(time.time() + 25)
        ps = [subprocess.Popen([sys.executable, __file__, str(i), "1" if numa else "0", "child"]) for i in range(n)]
        for p in ps: p.wait()
