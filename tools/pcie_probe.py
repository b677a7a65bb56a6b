#!/usr/bin/env python
"""Ceiling of the host<->device copies behind `e2e`: every rank copies the bytes one
bench step moves per GPU -- H2D 8*(nq+2nv) and D2H 8*nv bytes per state, 2^20 states -- between PINNED host
buffers and its GPU with plain cudaMemcpyAsync (one call per direction and step, torch `copy_`), all ranks
at the same time. Prints one JSON line: per-GPU and aggregate GB/s for H2D alone, D2H alone and both
directions together (two streams), CUDA-event timed, max over ranks.

    python tools/pcie_probe.py                                        # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/pcie_probe.py
"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, nq, nv = 1 << 20, 28, 27
    h_in = torch.empty(n * (nq + 2 * nv), dtype=torch.float64).pin_memory()
    h_out = torch.empty(n * nv, dtype=torch.float64).pin_memory()
    h_in.fill_(1.0)
    d_in = torch.empty_like(h_in, device=dev)
    d_out = torch.ones(n * nv, dtype=torch.float64, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    steps = 10

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(do_in, do_out):
        for _ in range(2):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e2 = torch.cuda.Event(enable_timing=True)
        e0.record()
        s_in.wait_event(e0)
        s_out.wait_event(e0)
        for _ in range(steps):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
        e1.record(s_in)
        e2.record(s_out)
        barrier()
        ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]) / steps

    bi, bo = h_in.numel() * 8, h_out.numel() * 8
    ms_in, ms_out, ms_both = timed(True, False), timed(False, True), timed(True, True)
    if rank == 0:
        line = {"probe": "pinned host <-> device, cudaMemcpyAsync, all ranks at once", "n_gpus": world,
                "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo,
                "h2d_ms": ms_in, "d2h_ms": ms_out, "both_ms": ms_both,
                "h2d_gbs_per_gpu": bi / ms_in * 1e-6, "d2h_gbs_per_gpu": bo / ms_out * 1e-6,
                "both_gbs_per_gpu": (bi + bo) / ms_both * 1e-6,
                "aggregate_both_gbs": world * (bi + bo) / ms_both * 1e-6,
                # a bench step through host buffers cannot be faster than its copies:
                "e2e_ceiling_states_per_s": world * n / (ms_both * 1e-3),
                "host_cpus": len(os.sched_getaffinity(0))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
