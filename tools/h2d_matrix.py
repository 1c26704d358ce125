"""Which GPUs share a host->device ceiling?  Concurrent pinned H2D copies on subsets of the box's GPUs, one thread and
one stream per GPU, 1 GiB per copy, 3 rounds; prints one JSON object {subset: {"total_gbs": .., "per_gpu_gbs": [..]}}.
Run on the multi-GPU box: python tools/h2d_matrix.py > gpurun_out/h2d_matrix.json"""
import json
import threading
import time

import torch

n = torch.cuda.device_count()
SIZE = 1 << 30
host = [torch.empty(SIZE, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(SIZE, dtype=torch.uint8, device=f"cuda:{i}") for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]


def run(subset, rounds=3):
    for i in subset:  # warm-up
        with torch.cuda.stream(streams[i]):
            dev[i].copy_(host[i], non_blocking=True)
    for i in subset:
        streams[i].synchronize()
    times = {}
    barrier = threading.Barrier(len(subset))

    def work(i):
        torch.cuda.set_device(i)
        barrier.wait()
        t0 = time.perf_counter()
        with torch.cuda.stream(streams[i]):
            for _ in range(rounds):
                dev[i].copy_(host[i], non_blocking=True)
        streams[i].synchronize()
        times[i] = time.perf_counter() - t0

    th = [threading.Thread(target=work, args=(i,)) for i in subset]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    return {"total_gbs": round(len(subset) * rounds * SIZE / wall / 1e9, 1),
            "per_gpu_gbs": [round(rounds * SIZE / times[i] / 1e9, 1) for i in subset]}


subsets = [[i] for i in range(n)]
if n >= 2:
    subsets += [[0, 1]]
if n >= 4:
    subsets += [[0, 2], [0, 3], [2, 3], [0, 1, 2, 3]]
if n >= 8:
    subsets += [[0, 4], [0, 7], [4, 5], [6, 7], [0, 1, 4, 5], [0, 2, 4, 6], [4, 5, 6, 7], list(range(8))]
out = {"gpus": n, "bytes_per_copy": SIZE, "subsets": {",".join(map(str, s)): run(s) for s in subsets}}
print(json.dumps(out))
