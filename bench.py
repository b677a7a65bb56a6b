#!/usr/bin/env python
"""Benchmark of the hot path: batched mj_inverse states/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one mjb_inverse pass over a resident batch of 2^20 synthetic humanoid states per GPU
(weak scaling: every rank owns its own 2^20-state shard, no data-path collective).

  value   device-resident throughput: states of all ranks / max-over-ranks CUDA-event time of the
          K timed steps (inputs already in HBM as structure-of-arrays)
  e2e     the same metric through the public C-ABI with HOST buffers: mjb_inverseHost (H2D from
          pinned memory + transposes, kernels, transpose + D2H of qfrc_inverse, pipelined in pieces
          over three streams), all inside the timed region
  roofline      HBM roofline of the step: algorithmic bytes (in + out per state) x states /
                CUDA-event time vs MEASURED_PEAKS.json HBM copy bandwidth, with the measured DRAM
                traffic of the step as `traffic` (the phase kernels are bound by the HBM traffic of
                the per-state intermediates)
  roofline_fp64 executed fp64 flops per state (ncu, profiles/flops_per_state.json) x states /
                CUDA-event time vs the DFMA peak measured in the same run
  kernels       per phase kernel: live CUDA-event time (events around every launch), share of the
                step, achieved fp64 TFLOP/s and DRAM GB/s against both peaks
  cpu_baseline  the reference's own mj_inverse looped over the host cores with its thread pool
                (oracle/_ref), on a bounded sample of the same states

`--impl reference` times only that CPU loop (rank 0) and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (golden model, z_range, description)
    "humanoid_contact_pyramidal": ("humanoid", (0.0, 1.5),
                                   "humanoid.xml, floor+self contacts, pyramidal cone"),
    "humanoid_contact_elliptic": ("humanoid_elliptic", (0.0, 1.5),
                                  "humanoid.xml, floor+self contacts, elliptic cone"),
    "humanoid_nocontact": ("humanoid_nocontact", (2.0, 3.0),
                           "humanoid.xml, mjDSBL_CONTACT"),
    "humanoids22": ("humanoids22", (0.0, 1.5), "22_humanoids.xml (nv=594)"),
    "arm26": ("arm26", (0.0, 1.5), "tendon_arm/arm26.xml (tendons, joint limits)"),
}
# the other BASELINE configs, measured device-resident after the headline (N = 1): config 2, 3b, one
# point of config 4's sweep, config 5
OTHER_CONFIGS = ["humanoid_nocontact", "humanoid_contact_elliptic", "arm26", "humanoids22"]
DEFAULT_WORKLOAD = "humanoid_contact_pyramidal"
BATCH = {"humanoids22": 1 << 15}
DEFAULT_BATCH = 1 << 20

# FP64 flops executed per state by the fused kernel (dadd + dmul + 2*dfma, thread-level, ncu
# smsp__sass_thread_inst_executed_op_d{add,mul,fma}_pred_on over a 2^20-state launch divided by
# the states); see profiles/ and DESIGN.md. Keyed by workload.
FLOPS_PER_STATE_FILE = os.path.join(ROOT, "profiles", "flops_per_state.json")


def load_flops_per_state(workload):
    """(fp64 flops per state, measured DRAM bytes per state, per-kernel records) from ncu."""
    try:
        with open(FLOPS_PER_STATE_FILE) as f:
            rec = json.load(f)[workload]
        return float(rec["flops_per_state"]), rec.get("dram_bytes_per_state"), rec.get("kernels", {})
    except Exception:
        return None, None, {}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False

    def start(self):
        # NVML in-process (a sample every few ms, so even a 100 ms timed region is covered);
        # nvidia-smi -lms as the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                ("sw_power_cap", 0x4))
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                smax = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                flags = ["Active" if (r & b) else "Not Active" for _, b in bits]
                self.samples.append(", ".join([str(sm), str(smax)] + flags))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=2)
        elif not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); smax = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(workload, n, world, no_inertia=False):
    """`config` of the JSON line: the same dict in the mjb arm and in the reference arm (the
    reference arm's bounded sample is described in its `cpu_baseline.sample`)."""
    return {"workload": workload, "description": WORKLOADS[workload][2], "states_per_gpu": n,
            "outputs": "qfrc_inverse" + ("" if no_inertia else "+qM+qLD+qLDiagInv"),
            "l2": "inputs larger than L2 (hundreds of MB of states per step)",
            "parallelism": f"batch sharded over {world} GPU(s), no collective"}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_reference_model(golden_name):
    """The reference library's own mjModel for the CPU baseline (oracle/_ref, checker only)."""
    import gzip
    import tempfile
    from oracle import reflib
    path = os.path.join(ROOT, "tests", "golden", golden_name + ".mjb.gz")
    with tempfile.NamedTemporaryFile(suffix=".mjb", delete=False) as tf:
        tf.write(gzip.open(path, "rb").read())
    try:
        return reflib.Model.from_mjb(tf.name)
    finally:
        os.remove(tf.name)


def cpu_reference_rate(golden_name, z_range, nthread, target_seconds=10.0, first=0):
    """states/s of the reference mj_inverse loop on `nthread` host threads, bounded sample."""
    from mujoco_inversedynamicstest_b200.states import generate_states
    rm = load_reference_model(golden_name)
    probe = 2048
    qpos, qvel, qacc = generate_states(rm, probe, first=first, z_range=z_range)
    rm.inverse_batch(qpos, qvel, qacc, nthread=nthread)                 # warm-up
    _, t = rm.inverse_batch(qpos, qvel, qacc, nthread=nthread)
    rate = probe / max(t, 1e-9)
    n = int(min(max(rate * target_seconds, probe), 4 << 20))
    qpos, qvel, qacc = generate_states(rm, n, first=first, z_range=z_range)
    _, t = rm.inverse_batch(qpos, qvel, qacc, nthread=nthread)
    return n / t, n, t


def specialise(bd, generic):
    """Switch a batch to the kernels compiled for its model (NVRTC, cached next to libmjb.so); the
    generic kernels stay in use when that is not possible, and the line says which ran."""
    if generic:
        return "generic"
    try:
        info = bd.specialize()
        return "specialised (cubin %s)" % ("from cache" if info["from_cache"] else
                                           "compiled in %.0f s" % info["compile_seconds"])
    except Exception as exc:
        return "generic (not specialised: %s)" % str(exc).splitlines()[0][:120]


def device_rate(mjb, workload, local_rank, stream, generic, steps=5, warmup=3):
    """Short device-resident measurement of another BASELINE config (same timing rules)."""
    import torch
    from mujoco_inversedynamicstest_b200.states import generate_states
    golden_name, z_range, _ = WORKLOADS[workload]
    n = BATCH.get(workload, DEFAULT_BATCH)
    model = mjb.Model.from_mjb(os.path.join(ROOT, "tests", "golden", golden_name + ".mjb.gz"))
    qpos, qvel, qacc = generate_states(model, n, z_range=z_range)
    bd = mjb.BatchData(model, n, device=local_rank, outmask=mjb.OUT_INERTIA, stream=stream.cuda_stream)
    try:
        mode = specialise(bd, generic)
        bd.set_state(qpos, qvel, qacc)
        for _ in range(warmup):
            bd.inverse(sync=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            bd.inverse(sync=False)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        nq, nv = model.int("nq"), model.int("nv")
        alg = 8 * (nq + 2 * nv) + 8 * nv + 4 + 8 * (model.int("nM") + model.int("nC") + nv)
        return {"workload": workload, "states_per_gpu": n, "steps": steps, "warmup": warmup,
                "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "states/s",
                "kernel_mode": mode, "bytes_per_state": alg,
                "hbm_gbs_algorithmic": n * alg / (ms * 1e-3) * 1e-9}
    finally:
        bd.close()


def parity_sample(mjb, workload, local_rank, stream, generic, nstate=1 << 16, first=1 << 23):
    """`parity` of the JSON line: a sample of the workload's stream that is in no committed fixture,
    through the CUDA path and through the reference library (oracle/_ref, the checker): states whose
    ncon / nefc / nl / nf / ne, contact geom pairs or efc_type / efc_id / efc_state differ, and
    qfrc_inverse entries outside 1e-9 * |ref| + 1e-12 (element-wise)."""
    from mujoco_inversedynamicstest_b200.states import generate_states
    golden_name, z_range, _ = WORKLOADS[workload]
    rm = load_reference_model(golden_name)
    model = mjb.Model.from_mjb(os.path.join(ROOT, "tests", "golden", golden_name + ".mjb.gz"))
    nconmax, njmax = 64, 256
    qpos, qvel, qacc = generate_states(model, nstate, first=first, z_range=z_range)
    ref, _ = rm.inverse_batch(qpos, qvel, qacc, nthread=host_threads(), fields={
        "ncon": 1, "nefc": 1, "nl": 1, "nf": 1, "ne": 1, "contact_geom": nconmax,
        "efc_type": njmax, "efc_id": njmax, "efc_state": njmax})
    bd = mjb.BatchData(model, nstate, device=local_rank, stream=stream.cuda_stream, nconmax=nconmax, njmax=njmax,
                       outmask=mjb.OUT_COUNTS | mjb.OUT_CONTACT | mjb.OUT_EFC | mjb.OUT_INERTIA)
    try:
        specialise(bd, generic)
        bd.set_state(qpos, qvel, qacc)
        flagged = bd.inverse()
        cnt, efc = bd.counts(), bd.efc()
        bad = np.zeros(nstate, dtype=bool)
        for key in ("ncon", "nefc", "nl", "nf", "ne"):
            bad |= cnt[key] != ref[key]
        bad |= (bd.contacts()["geom"] != ref["contact_geom"]).any(axis=(1, 2))
        for key in ("type", "id", "state"):
            bad |= (efc[key] != ref["efc_" + key][:, :, 0]).any(axis=1)
        got = bd.qfrc_inverse()
        d = np.abs(got - ref["qfrc_inverse"])
        tol = 1e-12 + 1e-9 * np.abs(ref["qfrc_inverse"])
        smax = np.maximum(np.abs(ref["qfrc_inverse"]).max(axis=1, keepdims=True), 1.0)
        return {"states": nstate, "flagged": int(flagged), "discrete_mismatch": int(bad.sum()),
                "entries": int(got.size), "strict_viol": int((d > tol).sum()),
                "strict_worst_ratio": float((d / tol).max()),
                "worst_error_over_state_max": float((d / smax).max()),
                "contacts": int(ref["ncon"].sum()), "against": "oracle/_ref (reference library, live)"}
    finally:
        bd.close()


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU mj_inverse over the host cores (rank 0 only)."""
    if rank != 0:
        return
    golden_name, z_range, desc = WORKLOADS[args.workload]
    nthread = host_threads()
    from mujoco_inversedynamicstest_b200.states import generate_states
    rm = load_reference_model(golden_name)
    # bounded sample per step: ~2 s of CPU work
    probe = 2048
    qpos, qvel, qacc = generate_states(rm, probe, z_range=z_range)
    rm.inverse_batch(qpos, qvel, qacc, nthread=nthread)
    _, t = rm.inverse_batch(qpos, qvel, qacc, nthread=nthread)
    n = int(min(max(probe / max(t, 1e-9) * 2.0, probe), 1 << 20))
    qpos, qvel, qacc = generate_states(rm, n, z_range=z_range)
    for _ in range(args.warmup):
        rm.inverse_batch(qpos, qvel, qacc, nthread=nthread)
    total = 0.0
    for _ in range(args.steps):
        _, t = rm.inverse_batch(qpos, qvel, qacc, nthread=nthread)
        total += t
    value = n * args.steps / total
    sample = (f"{n} states/step x {args.steps} steps of the {args.workload} stream, reference mj_inverse "
              f"+ src/thread pool on {nthread} host threads")
    line = {
        "impl": "reference", "metric": "mj_inverse states/sec (humanoid, fp64)", "value": value,
        "unit": "states/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, args.batch or BATCH.get(args.workload, DEFAULT_BATCH),
                                  max(world, 1), args.no_inertia),
        "cpu_baseline": {"value": value, "unit": "states/s", "cores": nthread, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    # the contract is ONE JSON line on stdout: route everything else written to fd 1 by libraries
    # (NCCL version banner, ...) to stderr and keep the real stdout for the result line
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mjb", choices=["mjb", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="states per GPU per step")
    ap.add_argument("--no-inertia", action="store_true",
                    help="skip the mj_crb / mj_factorM outputs (qM, qLD, qLDiagInv)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--generic", action="store_true",
                    help="generic kernels only (default: the kernels specialised for the model, mjb_specialize)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short device-timed lines of the other BASELINE configs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "mjb" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import mujoco_inversedynamicstest_b200 as mjb
    from mujoco_inversedynamicstest_b200.states import generate_states

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmjb has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    golden_name, z_range, desc = WORKLOADS[args.workload]
    n = args.batch or BATCH.get(args.workload, DEFAULT_BATCH)
    model = mjb.Model.from_mjb(os.path.join(ROOT, "tests", "golden", golden_name + ".mjb.gz"))
    nq, nv = model.int("nq"), model.int("nv")
    outmask = 0 if args.no_inertia else mjb.OUT_INERTIA

    # this rank's shard of the global state stream: [rank*n, (rank+1)*n)
    from mujoco_inversedynamicstest_b200.shard import weak_shard
    first_state, _ = weak_shard(n, rank)
    qpos, qvel, qacc = generate_states(model, n, first=first_state, z_range=z_range)
    h_qpos = torch.from_numpy(qpos).pin_memory()
    h_qvel = torch.from_numpy(qvel).pin_memory()
    h_qacc = torch.from_numpy(qacc).pin_memory()
    h_out = torch.empty((n, nv), dtype=torch.float64).pin_memory()

    stream = torch.cuda.current_stream()
    bd = mjb.BatchData(model, n, device=local_rank, outmask=outmask, stream=stream.cuda_stream)
    kernel_mode = specialise(bd, args.generic)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (value) ----------------
    bd.set_state_ptr(n, h_qpos.data_ptr(), h_qvel.data_ptr(), h_qacc.data_ptr())
    for _ in range(args.warmup):
        bd.inverse(sync=False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = bd.kernel_launches()
    ev0.record(stream)
    for _ in range(args.steps):
        bd.inverse(sync=False)
    ev1.record(stream)
    gpu_launches = bd.kernel_launches() - launches0
    barrier()
    kernel_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    got = bd.qfrc_inverse()
    if not np.isfinite(got).all():
        raise SystemExit("non-finite qfrc_inverse in the benchmark batch")

    # ---------------- per-kernel times, live, CUDA events around every launch ----------------
    pt_steps = max(3, min(args.steps, 5))
    bd.phase_timing(True)
    for _ in range(pt_steps):
        bd.inverse(sync=False)
    phase_ms = {k: v / pt_steps for k, v in bd.phase_times().items()}
    bd.phase_timing(False)

    # ---------------- end-to-end through the C-ABI with host buffers ----------------
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        bd.inverse_host(n, h_qpos.data_ptr(), h_qvel.data_ptr(), h_qacc.data_ptr(), h_out.data_ptr())
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        bd.inverse_host(n, h_qpos.data_ptr(), h_qvel.data_ptr(), h_qacc.data_ptr(), h_out.data_ptr())
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    t = torch.tensor([kernel_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    kernel_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        total_states = n * world
        value = total_states * args.steps / (kernel_ms * 1e-3)
        e2e_value = total_states * e2e_steps / (e2e_ms * 1e-3)
        peaks, peak_src = measured_peaks()
        per_gpu_rate = n * args.steps / (kernel_ms * 1e-3)
        fp64_peak = mjb.fp64_peak_tflops(local_rank)
        fps, dram_per_state, kcounts = load_flops_per_state(args.workload)
        if args.no_inertia:
            fps, dram_per_state, kcounts = None, None, {}   # the frozen counts include the inertia kernel
        alg_bytes = 8 * (nq + 2 * nv) + 8 * nv + 4
        if not args.no_inertia:
            alg_bytes += 8 * (model.int("nM") + model.int("nC") + nv)
        roof_fp64 = {
            "bound": "fp64", "achieved": (per_gpu_rate * fps * 1e-12) if fps else None,
            "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (per_gpu_rate * fps * 1e-12 / fp64_peak) if (fps and fp64_peak > 0) else None,
            "traffic": (dram_per_state * n) if dram_per_state else None,
            "traffic_unit": "DRAM bytes per step per GPU (ncu dram__bytes_read+write, profiles/)",
            "flops_per_state": fps,
            "kernel": "the step = smooth + inertia + contact_scan + contact + backward kernels",
            "peak_source": "DFMA probe measured in this run (mjb_fp64PeakTflops)"}
        # per kernel: live CUDA-event time of this run x the frozen ncu counts (flops, DRAM bytes)
        kernels = []
        for name, ms in phase_ms.items():
            if ms <= 0:
                continue
            kc = kcounts.get(name, {})
            fl, by = kc.get("flops"), kc.get("dram_bytes")
            kernels.append({
                "kernel": ("contact_narrow+index+rows kernels" if name == "contact" else name + "_kernel"),
                "ms_per_step": ms,
                "share": ms / max(sum(phase_ms.values()), 1e-12),
                "fp64_tflops": (fl * n / (ms * 1e-3) * 1e-12) if fl else None,
                "fp64_frac": (fl * n / (ms * 1e-3) * 1e-12 / fp64_peak) if (fl and fp64_peak > 0) else None,
                "dram_gbs": (by * n / (ms * 1e-3) * 1e-9) if by else None,
                "hbm_frac": (by * n / (ms * 1e-3) * 1e-9 / peaks["hbm_gbs"]) if by else None})
        kernels.sort(key=lambda k: -k["ms_per_step"])
        roof_hbm = {
            "bound": "hbm", "achieved": per_gpu_rate * alg_bytes * 1e-9, "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": per_gpu_rate * alg_bytes * 1e-9 / peaks["hbm_gbs"],
            "traffic": (dram_per_state * n) if dram_per_state else None,
            "traffic_unit": "DRAM bytes per step per GPU (ncu dram__bytes_read+write, profiles/)",
            "bytes_per_state": alg_bytes, "peak_source": peak_src,
            "kernel": "the whole step (all phase kernels; per-kernel figures under `kernels`)"}
        line = {
            "metric": "mj_inverse states/sec (humanoid, fp64)", "value": value, "unit": "states/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, n, world, args.no_inertia),
            "kernel_mode": kernel_mode,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "states/s",
                    "h2d_bytes_per_step": n * (nq + 2 * nv) * 8, "d2h_bytes_per_step": n * nv * 8,
                    "steps": e2e_steps},
            "gpu_launches": gpu_launches,
            "roofline": roof_hbm,
            "roofline_fp64": roof_fp64,
            "kernels": kernels,
            "kernels_note": "ms: CUDA events around every launch in this run (mjb_phaseTiming); flops "
                            "and DRAM bytes per state: ncu counts frozen in profiles/flops_per_state.json; "
                            "hbm_frac is MEASURED DRAM traffic / time vs the HBM copy peak",
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                nthread = host_threads()
                rate, ns, secs = cpu_reference_rate(golden_name, z_range, nthread)
                line["cpu_baseline"] = {
                    "value": rate, "unit": "states/s", "cores": nthread, "kind": "reference",
                    "sample": f"first {ns} states of the same stream, {secs:.1f} s, "
                              "reference mj_inverse + src/thread pool"}
            except Exception as exc:  # the checker library is absent: say so, do not fake it
                line["cpu_baseline"] = {"value": None, "unit": "states/s", "cores": 0,
                                        "kind": "reference", "sample": f"unavailable: {exc}"}
        if world == 1 and not args.no_other_configs:
            # not part of the timed regions above: parity of a fresh sample against the reference
            # library, and short device-resident lines of the other BASELINE configs
            bd.close()
            try:
                line["parity"] = parity_sample(mjb, args.workload, local_rank, stream, args.generic)
            except Exception as exc:
                line["parity"] = {"unavailable": str(exc)[:200]}
            line["other_configs"] = []
            for wl in OTHER_CONFIGS:
                if wl == args.workload:
                    continue
                try:
                    line["other_configs"].append(device_rate(mjb, wl, local_rank, stream, args.generic))
                except Exception as exc:
                    line["other_configs"].append({"workload": wl, "unavailable": str(exc)[:200]})
        print(json.dumps(line), flush=True)

    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
