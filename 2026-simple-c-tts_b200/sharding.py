"""Utterance sharding across the GPUs of one box.

Utterances are independent (no state crosses ctts_synthesize calls except the
read-only voice DB), so a batch is partitioned by utterance with no data-path
collective: every rank holds a replica of the PCM pool, runs its shard through
its own ctts_gpu context, and the host gathers the output buffers.
"""
from __future__ import annotations

import heapq

import numpy as np


def utterance_costs(pre_bound: np.ndarray, speeds: np.ndarray, stretch_weight: float = 8.0) -> np.ndarray:
    """Relative cost per utterance: samples assembled, weighted up when WSOLA runs (speed != 1)."""
    c = pre_bound.astype(np.float64)
    s = np.asarray(speeds, dtype=np.float32)
    return np.where(s != np.float32(1.0), c * stretch_weight, c)


def shard_indices(costs: np.ndarray, world: int) -> list[np.ndarray]:
    """Greedy longest-processing-time partition; deterministic; every index appears exactly once.

    Within a shard the indices are returned in ascending order so a rank's outputs keep
    the batch order."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    heap = [(0.0, r) for r in range(world)]
    heapq.heapify(heap)
    shards: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        load, r = heapq.heappop(heap)
        shards[r].append(int(i))
        heapq.heappush(heap, (load + float(costs[i]), r))
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def gather_order(shards: list[np.ndarray]) -> np.ndarray:
    """Permutation that maps the concatenation of per-rank outputs back to batch order."""
    cat = np.concatenate(shards) if shards else np.zeros(0, np.int64)
    inv = np.empty(len(cat), dtype=np.int64)
    inv[cat] = np.arange(len(cat))
    return inv
