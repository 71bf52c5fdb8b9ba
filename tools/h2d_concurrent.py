#!/usr/bin/env python
"""Pinned-host -> device bandwidth with 1 / 2 / 4 / 8 GPUs copying AT THE SAME TIME (one process per GPU, as bench.py's
end-to-end arm runs): names the resource that caps end-to-end scaling.  Each rank copies 8 x 19 MB + 8 x 0.3 MB (one
bench step's inputs) 20 times from its own pinned buffers; all ranks start together (gloo barrier).  Prints one JSON
object: per-N aggregate and per-GPU GB/s, plus the box topology (NUMA nodes, CPU count, `nvidia-smi topo -m`).
  python tools/h2d_concurrent.py            (uses every visible GPU)"""
import json
import os
import subprocess
import sys
import time

import torch
import torch.multiprocessing as mp


def worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    T, D = 2325, 2048
    hx = [torch.empty(T, D).pin_memory() for _ in range(8)]
    hl = [torch.empty(T, 132, dtype=torch.uint8).pin_memory() for _ in range(8)]
    dx = torch.empty(8 * T, D, device=dev)
    dl = torch.empty(8 * T, 132, dtype=torch.uint8, device=dev)
    nbytes = 8 * T * (D * 4 + 132)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]

    def step():
        for i in range(8):
            with torch.cuda.stream(streams[i & 1]):
                dx[i * T:(i + 1) * T].copy_(hx[i], non_blocking=True)
            with torch.cuda.stream(streams[(i + 1) & 1]):
                dl[i * T:(i + 1) * T].copy_(hl[i], non_blocking=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    res = {}
    for active in (1, 2, 4, 8):
        if active > world:
            break
        torch.distributed.barrier()
        dt = 0.0
        if rank < active:
            t0 = time.perf_counter()
            for _ in range(20):
                step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        torch.distributed.barrier()
        res[active] = nbytes * 20 / dt / 1e9 if dt > 0 else None
    q.put((rank, res))
    torch.distributed.destroy_process_group()


def main():
    world = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, 29611, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get() for _ in procs]
    for p in procs:
        p.join()
    table = {}
    for active in (1, 2, 4, 8):
        vals = [res[active] for _, res in sorted(out) if res.get(active)]
        if vals:
            table[str(active)] = {"aggregate_GBps": round(sum(vals), 1), "per_gpu_GBps": [round(v, 1) for v in vals]}

    def sh(cmd):
        try:
            return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
        except Exception as e:  # noqa: BLE001
            return repr(e)

    info = {"gpus": world, "cpus": os.cpu_count(), "numa_nodes": sh("ls -d /sys/devices/system/node/node* | wc -l"),
            "cpu_model": sh("lscpu | grep 'Model name' | head -1"), "virtualization": sh("lscpu | grep -i 'hypervisor vendor' | head -1"),
            "mem_total": sh("grep MemTotal /proc/meminfo"), "topo": sh("nvidia-smi topo -m | head -14")}
    print(json.dumps({"concurrent_h2d": table, "box": info}))


if __name__ == "__main__":
    sys.exit(main())
