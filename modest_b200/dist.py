"""Sharding and the single collective of the seed-label path.

Scans are independent, so W ranks (one process per GPU, launched by torchrun) take the W
contiguous `np.array_split` shards the reference's `total_part` / `part` keys already define
(pre_compute_pp_score.py:114-116, generate_mask.py:35-37, gen_label_files.py:36-38); no point
data ever crosses ranks.  The only exchange is collating the per-scan label blobs: one
all-gather of lengths, one of the padded payload (NCCL on CUDA tensors, gloo on CPU tensors).
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as td


def world_size() -> int:
    return td.get_world_size() if td.is_available() and td.is_initialized() else int(os.environ.get("WORLD_SIZE", 1))


def rank() -> int:
    return td.get_rank() if td.is_available() and td.is_initialized() else int(os.environ.get("RANK", 0))


def resolve_parts(total_part, part):
    """`total_part`/`part` as configured; when left at the defaults (1, 0) under torchrun they
    follow WORLD_SIZE / RANK."""
    total_part, part = int(total_part), int(part)
    if total_part == 1 and part == 0 and int(os.environ.get("WORLD_SIZE", 1)) > 1:
        return int(os.environ["WORLD_SIZE"]), int(os.environ.get("RANK", 0))
    return total_part, part


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for one process)."""
    if int(os.environ.get("WORLD_SIZE", 1)) <= 1 or td.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    td.init_process_group(backend=backend)


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(device_index=None, sysfs="/sys/bus/pci/devices"):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging
    buffers allocated afterwards (first touch) are local to that GPU's PCIe root: a staging buffer
    on the other socket halves or thirds the host->device rate of the streaming engine.  Returns
    the CPU set chosen, or None when the topology is not exposed (then nothing is changed)."""
    try:
        if not torch.cuda.is_available() or not hasattr(os, "sched_setaffinity"):
            return None
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        pr = torch.cuda.get_device_properties(idx)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        with open(os.path.join(sysfs, bdf, "local_cpulist")) as fh:
            local = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        chosen = local & allowed
        if not chosen or chosen == allowed:
            return None
        os.sched_setaffinity(0, chosen)
        return chosen
    except (OSError, AttributeError, ValueError):
        return None


def shard(items, world=None, r=None):
    world = world_size() if world is None else world
    r = rank() if r is None else r
    return list(np.array_split(np.asarray(list(items)), world)[r]) if world > 1 else list(items)


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if td.get_backend() == "nccl" else torch.device("cpu")


def gather_blobs(local: dict) -> dict:
    """{scan id: bytes} of this rank -> the union over all ranks, on every rank."""
    if not (td.is_available() and td.is_initialized()) or td.get_world_size() == 1:
        return dict(local)
    dev, W = _device(), td.get_world_size()
    ids = np.array(sorted(local), dtype=np.int64)
    lens = np.array([len(local[i]) for i in ids], dtype=np.int64)
    payload = b"".join(local[i] for i in ids)
    head = torch.tensor([len(ids), len(payload)], dtype=torch.int64, device=dev)
    heads = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(W)]
    td.all_gather(heads, head)
    heads = torch.stack(heads).cpu().numpy()
    max_n, max_b = int(heads[:, 0].max()), int(heads[:, 1].max())
    meta = torch.zeros(2 * max_n, dtype=torch.int64, device=dev)
    meta[:len(ids)] = torch.from_numpy(ids).to(dev)
    meta[max_n:max_n + len(ids)] = torch.from_numpy(lens).to(dev)
    buf = torch.zeros(max(max_b, 1), dtype=torch.uint8, device=dev)
    if payload:
        buf[:len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    metas = [torch.zeros_like(meta) for _ in range(W)]
    bufs = [torch.zeros_like(buf) for _ in range(W)]
    td.all_gather(metas, meta)
    td.all_gather(bufs, buf)
    out = {}
    for r in range(W):
        n = int(heads[r, 0])
        m = metas[r].cpu().numpy()
        data = bufs[r].cpu().numpy().tobytes()
        pos = 0
        for k in range(n):
            ln = int(m[max_n + k])
            out[int(m[k])] = data[pos:pos + ln]
            pos += ln
    return out
