"""Synthetic workloads: seeded, uniformly sampled joint states and IK targets.

Counter-based generator (splitmix64 on seed + element counter), so that any slice of a
batch can be generated independently on any rank and the same arrays can be handed to the
GPU path and to the CPU reference:
  q_j  ~ U[q_min_j, q_max_j]  for unconstrained, uncoupled joints,
  q_j  = q0_j                 for constrained joints,
  q_j  = q0_j                 for coupled (slave) joints — the libraries overwrite them from
                              the master exactly as RcsGraph_setState does,
  qd_j ~ U[-1, 1]             (0 for constrained joints).
"""
from __future__ import annotations

import numpy as np

SEED = 0x5243535F42323030  # "RCS_B200"


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(seed: int, first: int, count: int, width: int, stream: int = 0) -> np.ndarray:
    """count x width doubles in [0, 1); element (i, k) depends only on (seed, stream, first+i, k)."""
    with np.errstate(over="ignore"):
        idx = (np.arange(first, first + count, dtype=np.uint64)[:, None] * np.uint64(width)
               + np.arange(width, dtype=np.uint64)[None, :])
        key = np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0xD1342543DE82EF95))
        bits = _splitmix64(idx ^ key)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def sample_states(layout: dict, n: int, first: int = 0, seed: int = SEED, with_velocity=False,
                  margin: float = 0.0):
    """Full-state (n x dof) q (and qd) for the graph described by a layout dict.

    margin in [0, 0.5): fraction of the joint range kept clear at both limits.
    """
    joints = layout["joints"]
    dof = layout["dof"]
    q0 = np.array([j["q0"] for j in joints])
    lo = np.array([j["q_min"] for j in joints])
    hi = np.array([j["q_max"] for j in joints])
    free = np.array([(not j["constrained"]) and j["coupledTo"] < 0 for j in joints])
    u = uniform01(seed, first, n, dof, stream=0)
    span = hi - lo
    q = lo + (margin + (1.0 - 2.0 * margin) * u) * span
    q = np.where(free[None, :], q, q0[None, :])
    if not with_velocity:
        return np.ascontiguousarray(q)
    v = 2.0 * uniform01(seed, first, n, dof, stream=1) - 1.0
    unconstrained = np.array([not j["constrained"] for j in joints])
    qd = np.where(unconstrained[None, :], v, 0.0)
    return np.ascontiguousarray(q), np.ascontiguousarray(qd)
