#!/usr/bin/env python3
"""NVLink bytes each collective kernel really moves, from the link counters of the driver (`nvidia-smi nvlink -gt d`: data Tx / Rx KiB
per link), next to its algorithmic bytes and the device time: achieved fraction of the 900 GB/s per-direction link roofline.
Run under torchrun; every rank reads the counters of its own GPU before and after K back-to-back launches.
(ncu cannot profile these kernels across ranks: it serialises and replays kernels, and a kernel that waits for its peers is not replayable.)"""
import json
import os
import re
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucc_b200.dist import Communicator, init_distributed  # noqa: E402

rank, N, lrank = init_distributed("cpu:gloo,cuda:nccl")
dev = torch.device("cuda", torch.cuda.current_device())
gpu_index = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[lrank] if os.environ.get("CUDA_VISIBLE_DEVICES") else str(lrank)
stream = torch.cuda.Stream()
K = int(os.environ.get("NVL_ITERS", "20"))


def counters():
    """(tx_bytes, rx_bytes) summed over the links of this rank's GPU, or None"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", gpu_index], capture_output=True, text=True, timeout=20).stdout
    except Exception:  # noqa: BLE001
        return None
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    if not re.search(r"Data Tx", out):
        return None
    return tx * 1024, rx * 1024


def measure(name, comm, make_req, alg_tx, alg_rx):
    """make_req() -> a fresh request; alg_tx / alg_rx = bytes the algorithm has to send / receive per GPU and launch"""
    reqs = [make_req() for _ in range(3 + K)]
    with torch.cuda.stream(stream):
        for r in reqs[:3]:
            r.post_on_stream(stream)
        for r in reqs[:3]:
            r.wait()
        torch.cuda.synchronize(); dist.barrier()
        c0 = counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for r in reqs[3:]:
            r.post_on_stream(stream)
        e1.record(stream)
        for r in reqs[3:]:
            r.wait()
        torch.cuda.synchronize(); dist.barrier()
        c1 = counters()
    for r in reqs:
        r.finalize()
    us = e0.elapsed_time(e1) * 1e3 / K
    row = {"kernel": comm.request_info_last(), "case": name, "rank": rank, "us": round(us, 1), "alg_tx_MB": round(alg_tx / 1e6, 2), "alg_rx_MB": round(alg_rx / 1e6, 2)}
    if c0 and c1:
        tx, rx = (c1[0] - c0[0]) / K, (c1[1] - c0[1]) / K
        row.update({"nvlink_tx_MB": round(tx / 1e6, 2), "nvlink_rx_MB": round(rx / 1e6, 2), "tx_over_alg": round(tx / alg_tx, 3) if alg_tx else None,
                    "tx_GBps": round(tx / us / 1e3, 1), "rx_GBps": round(rx / us / 1e3, 1), "frac_of_900_tx": round(tx / us / 1e3 / 900, 3)})
    else:
        row["nvlink_counters"] = "unavailable (nvidia-smi nvlink -gt d)"
    rows = [None] * N
    dist.all_gather_object(rows, row)
    if rank == 0:
        print(json.dumps({"case": name, "per_rank": rows}), flush=True)


def force(alg):
    return [("tl/nvl", "TUNE", alg), ("tl/nvl", "ALLREDUCE_ONESHOT_THRESH", "0")]


S = int(os.environ.get("NVL_BYTES", str(256 << 20)))
cnt = S // 4
f = (N - 1) / N
for name, mods, symm in (("allreduce twoshot zero-copy", force("allreduce:cuda:inf:@twoshot"), None), ("allreduce nvls staged", force("allreduce:cuda:inf:@nvls"), None),
                         ("allreduce nvls symmetric in-place", [], "2G"), ("allreduce oneshot 256K", [], None)):
    try:
        comm = Communicator(ctx_modify=mods, symm_size=symm)
    except Exception as e:  # noqa: BLE001
        if rank == 0:
            print(json.dumps({"case": name, "error": str(e)}), flush=True)
        continue
    if "oneshot" in name:
        c2 = (256 << 10) // 4
        src, dst = torch.ones(c2, device=dev), torch.empty(c2, device=dev)
        measure(name, comm, lambda: comm.allreduce_init(src, dst), c2 * 4 * (N - 1), c2 * 4 * (N - 1))
    else:
        if symm and comm.symm_region() is not None:
            src, dst = comm.symm_empty(cnt), comm.symm_empty(cnt)
            src.fill_(1)
        else:
            src, dst = torch.ones(cnt, device=dev), torch.empty(cnt, device=dev)
        torch.cuda.synchronize()
        # two-shot: every GPU sends its peers' slices of its vector (reduce-scatter reads) + the reduced slice to N-1 peers; in the switch: S*(N-1)/N out + S/N multicast
        alg = 2 * S * f if "twoshot" in name else S * f + S / N
        measure(name, comm, lambda: comm.allreduce_init(src, dst), alg, alg)
    del src, dst
    comm.destroy()
for name, tune in (("allgather push (tma)", "allgather:cuda:inf:@push"), ("alltoall push (tma)", "alltoall:cuda:inf:@push"), ("alltoall copy engine", "alltoall:cuda:inf:@ce")):
    comm = Communicator(ctx_modify=[("tl/nvl", "TUNE", tune)])
    blk = cnt // N
    if "allgather" in name:
        src, dst = torch.ones(blk, device=dev), torch.empty(blk * N, device=dev)
        measure(name, comm, lambda: comm.coll_init("allgather", src, dst), blk * 4 * (N - 1), blk * 4 * (N - 1))
    else:
        src, dst = torch.ones(blk * N, device=dev), torch.empty(blk * N, device=dev)
        measure(name, comm, lambda: comm.coll_init("alltoall", src, dst), blk * 4 * (N - 1), blk * 4 * (N - 1))
    del src, dst
    comm.destroy()
dist.destroy_process_group()
