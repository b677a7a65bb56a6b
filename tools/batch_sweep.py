"""BASELINE.json config 4: "tendons, equality constraints and joint limits through mj_invConstraint,
batch sweep 1K-4M" -- states/s of mjb_inverse over batch sizes 2^10 .. 2^22 for the config-4 models
(arm26, slider_crank with contacts disabled, weld, connect, the zoo scene), device-resident and end to
end through host buffers, next to the reference's threaded mj_inverse loop on the host cores.

    python tools/batch_sweep.py > gpurun_out/config4_sweep.jsonl     (on the GPU box)

One JSON line per (model, batch). Inputs of the small batches fit the 126 MB L2 between steps; the
line says so (`inputs_bytes`). Timing: CUDA events on the launching stream, 3 warm-up steps.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import mujoco_inversedynamicstest_b200 as mjb  # noqa: E402
from mujoco_inversedynamicstest_b200.states import generate_states  # noqa: E402

MODELS = {"arm26": (0.0, 1.5), "slider_crank_nocontact": (0.0, 1.5), "weld": (0.0, 1.5),
          "connect": (0.0, 1.5), "zoo": (0.0, 0.6)}
BATCHES = [1 << k for k in range(10, 23, 2)]


def main():
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream()
    nthread = bench.host_threads()
    for name, zr in MODELS.items():
        model = mjb.Model.from_mjb(os.path.join(ROOT, "tests", "golden", name + ".mjb.gz"))
        nq, nv = model.int("nq"), model.int("nv")
        cpu_rate, cpu_n, cpu_t = bench.cpu_reference_rate(name, zr, nthread, target_seconds=3.0)
        for n in BATCHES:
            qpos, qvel, qacc = generate_states(model, n, z_range=zr)
            bd = mjb.BatchData(model, n, device=0, outmask=0, stream=stream.cuda_stream)
            bd.set_state(qpos, qvel, qacc)
            steps = max(5, min(200, (1 << 24) // n))
            for _ in range(3):
                bd.inverse(sync=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                bd.inverse(sync=False)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            h = [torch.from_numpy(a).pin_memory() for a in (qpos, qvel, qacc)]
            h_out = torch.empty((n, nv), dtype=torch.float64).pin_memory()
            esteps = max(3, min(50, (1 << 22) // n))
            for _ in range(2):
                bd.inverse_host(n, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h_out.data_ptr())
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(esteps):
                bd.inverse_host(n, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h_out.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            ems = e0.elapsed_time(e1) / esteps
            assert np.isfinite(h_out.numpy()).all()
            print(json.dumps({
                "config": "BASELINE config 4 batch sweep", "model": name, "nv": nv, "batch": n,
                "ms_per_step": ms, "states_per_s": n / (ms * 1e-3), "steps": steps,
                "e2e_states_per_s": n / (ems * 1e-3), "e2e_steps": esteps,
                "inputs_bytes": 8 * n * (nq + 2 * nv),
                "cpu_reference": {"states_per_s": cpu_rate, "cores": nthread, "sample": cpu_n, "seconds": cpu_t}}),
                flush=True)
            bd.close()


if __name__ == "__main__":
    main()
