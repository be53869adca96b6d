"""Concurrent host->device bandwidth of N GPUs of one box, with the NUMA placement of every pinned buffer.

    python tools/probes/h2d_concurrent.py [--mb 139] [--gpus 1,2,4,8]

For each N: N processes, one per GPU, each copying its own pinned buffer to its own GPU with bare cudaMemcpyAsync
(torch copy_ from pinned memory, non_blocking) for ~1 s, all started together; prints per-GPU and aggregate GB/s.
Done twice: buffers placed by first touch (where the process runs), and -- when the kernel lets this cgroup allocate
there -- bound to the GPU's own NUMA node with set_mempolicy(MPOL_BIND) before the pinned allocation.  The point: tell
a platform ceiling (PCIe switch / socket interconnect shared by several GPUs) from a placement mistake.
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import time


def gpu_numa_node(index: int) -> int:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    path = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
    try:
        return int(open(path).read())
    except Exception:
        return -1


def allowed(name: str) -> str:
    for line in open("/proc/self/status"):
        if line.startswith(name):
            return line.split(":")[1].strip()
    return "?"


def bind_memory(node: int) -> bool:
    """set_mempolicy(MPOL_BIND, {node}): later allocations of this process come from `node` only."""
    if node < 0:
        return False
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    SYS_set_mempolicy, MPOL_BIND = 238, 2                      # x86-64
    return libc.syscall(SYS_set_mempolicy, MPOL_BIND, ctypes.byref(mask), ctypes.c_ulong(64)) == 0


def page_node(addr: int) -> int:
    """NUMA node that backs the page at `addr` (move_pages with a NULL node list queries)."""
    libc = ctypes.CDLL(None, use_errno=True)
    pages = (ctypes.c_void_p * 1)(addr)
    status = (ctypes.c_int * 1)(-1)
    SYS_move_pages = 279
    rc = libc.syscall(SYS_move_pages, 0, ctypes.c_ulong(1), pages, None, status, 0)
    return int(status[0]) if rc == 0 else -1


def worker(rank, n, mb, bind, barrier, out):
    import torch
    node = gpu_numa_node(rank)
    bound = bind_memory(node) if bind else False
    torch.cuda.set_device(rank)
    nbytes = mb << 20
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.fill_(rank)                                              # first touch
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier.wait()
    reps = 0
    a.record()
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.0:
        for _ in range(4):
            d.copy_(h, non_blocking=True)
        reps += 4
        torch.cuda.current_stream().synchronize()
    b.record()
    torch.cuda.synchronize()
    out.put({"gpu": rank, "gpu_node": node, "buffer_node": page_node(h.data_ptr()), "bound": bound,
             "cpus": sorted(os.sched_getaffinity(0))[:1] + [len(os.sched_getaffinity(0))],
             "gbs": nbytes * reps / a.elapsed_time(b) / 1e6})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=139)             # the bench's 1080p sequence per step
    ap.add_argument("--gpus", default="1,2,4,8")
    args = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    print(json.dumps({"gpus_visible": have, "cpus_allowed": allowed("Cpus_allowed_list"), "mems_allowed": allowed("Mems_allowed_list"),
                      "gpu_numa_nodes": [gpu_numa_node(i) for i in range(have)]}))
    ctx = mp.get_context("spawn")
    for bind in (False, True):
        for n in [int(x) for x in args.gpus.split(",") if int(x) <= have]:
            barrier, out = ctx.Barrier(n), ctx.Queue()
            ps = [ctx.Process(target=worker, args=(r, n, args.mb, bind, barrier, out)) for r in range(n)]
            for p in ps:
                p.start()
            rows = sorted((out.get() for _ in ps), key=lambda r: r["gpu"])
            for p in ps:
                p.join()
            print(json.dumps({"n": n, "bind_to_gpu_node": bind, "aggregate_gbs": round(sum(r["gbs"] for r in rows), 1),
                              "per_gpu_gbs": [round(r["gbs"], 1) for r in rows], "buffer_nodes": [r["buffer_node"] for r in rows],
                              "bound_ok": [r["bound"] for r in rows]}))


if __name__ == "__main__":
    main()
