"""Deterministic synthetic (qpos, qvel, qacc) states shared by the CPU reference, the tests and the
benchmark.

Counter-based: every number is a pure function of (seed, state index, field index), so any shard
of a batch can be regenerated independently on any rank.

    hinge/slide qpos : uniform in [lo - 0.1 w, hi + 0.1 w] when the joint is limited (w = hi - lo),
                       so both active and inactive limits occur; else qpos0 + U(-1, 1)
    free joint       : x, y ~ U(-1, 1); z ~ U(zlo, zhi); quaternion = normalised N(0,1)^4
    ball joint       : quaternion = normalised N(0,1)^4
    qvel ~ U(-1, 1), qacc ~ U(-10, 10)
"""
import numpy as np

SEED = 20250331

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _uniform(seed, idx, field):
    """U[0,1) for state indices idx (uint64 array) and integer field id."""
    with np.errstate(over="ignore"):
        key = _splitmix(np.uint64(seed) ^ (np.uint64(field) * np.uint64(0xD6E8FEB86659FD93)))
        x = _splitmix(key ^ (idx * np.uint64(0x9E3779B97F4A7C15)))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _normal(seed, idx, field):
    u1 = _uniform(seed, idx, 2 * field + 1000003)
    u2 = _uniform(seed, idx, 2 * field + 1000004)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def generate_states(model, n, first=0, seed=SEED, z_range=(0.0, 1.5)):
    """model: object with .int(name) and .array(name) (the package Model or the test-side reference wrapper).

    Returns qpos [n, nq], qvel [n, nv], qacc [n, nv] for state indices first .. first+n-1."""
    nq, nv, njnt = model.int("nq"), model.int("nv"), model.int("njnt")
    jnt_type = model.array("jnt_type").ravel()
    jnt_qposadr = model.array("jnt_qposadr").ravel()
    jnt_limited = model.array("jnt_limited").ravel()
    jnt_range = model.array("jnt_range").reshape(-1, 2)
    qpos0 = model.array("qpos0").ravel()
    idx = np.arange(first, first + n, dtype=np.uint64)

    qpos = np.empty((n, nq))
    for j in range(njnt):
        t, a = int(jnt_type[j]), int(jnt_qposadr[j])
        if t == 0:      # free
            qpos[:, a + 0] = 2 * _uniform(seed, idx, a + 0) - 1
            qpos[:, a + 1] = 2 * _uniform(seed, idx, a + 1) - 1
            qpos[:, a + 2] = z_range[0] + (z_range[1] - z_range[0]) * _uniform(seed, idx, a + 2)
            q = np.stack([_normal(seed, idx, a + 3 + k) for k in range(4)], axis=1)
            qpos[:, a + 3:a + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
        elif t == 1:    # ball
            q = np.stack([_normal(seed, idx, a + k) for k in range(4)], axis=1)
            qpos[:, a:a + 4] = q / np.linalg.norm(q, axis=1, keepdims=True)
        else:           # slide, hinge
            u = _uniform(seed, idx, a)
            if jnt_limited[j]:
                lo, hi = jnt_range[j]
                w = hi - lo
                qpos[:, a] = (lo - 0.1 * w) + (1.2 * w) * u
            else:
                qpos[:, a] = qpos0[a] + (2 * u - 1)
    qvel = np.empty((n, nv))
    qacc = np.empty((n, nv))
    for i in range(nv):
        qvel[:, i] = 2 * _uniform(seed, idx, 100000 + i) - 1
        qacc[:, i] = 20 * _uniform(seed, idx, 200000 + i) - 10
    return qpos, qvel, qacc


def near_default_states(model, n, scale=0.05, seed=SEED):
    """State 0 = the model's default state (qpos0, zero velocity and acceleration), as the
    reference's own tests evaluate it; states 1 .. n-1 = qpos0 perturbed by `scale` (positions and
    scalar joints +- scale, quaternions re-normalised after a perturbation of N(0, scale)), with
    qvel ~ U(-1, 1) and qacc ~ U(-10, 10). Used for the reference's edge-case models, whose contact
    structure (stacked boxes, touching spheres) only exists near the default pose."""
    nq, nv, njnt = model.int("nq"), model.int("nv"), model.int("njnt")
    jnt_type = model.array("jnt_type").ravel()
    jnt_qposadr = model.array("jnt_qposadr").ravel()
    qpos0 = model.array("qpos0").ravel()
    idx = np.arange(0, n, dtype=np.uint64)
    qpos = np.tile(qpos0[None, :], (n, 1)) if nq else np.empty((n, 0))
    for j in range(njnt):
        t, a = int(jnt_type[j]), int(jnt_qposadr[j])
        if t == 0:
            for k in range(3):
                qpos[:, a + k] += scale * (2 * _uniform(seed, idx, 300000 + a + k) - 1)
            a += 3
        if t in (0, 1):
            q = qpos[:, a:a + 4] + scale * np.stack([_normal(seed, idx, 300000 + a + k) for k in range(4)], axis=1)
            qpos[:, a:a + 4] = q / np.linalg.norm(q, axis=1, keepdims=True)
        else:
            qpos[:, a] += scale * (2 * _uniform(seed, idx, 300000 + a) - 1)
    qvel = np.empty((n, nv))
    qacc = np.empty((n, nv))
    for i in range(nv):
        qvel[:, i] = 2 * _uniform(seed, idx, 100000 + i) - 1
        qacc[:, i] = 20 * _uniform(seed, idx, 200000 + i) - 10
    if n and nq:
        qpos[0] = qpos0
    if n:
        qvel[0] = 0
        qacc[0] = 0
    return qpos, qvel, qacc
