"""Host-side gather of a batch sharded over the GPUs of one box (BASELINE configs[3]).

There is no data-path collective: utterances are independent, every rank (one process per GPU) synthesises
its shard and its device->host copies land DIRECTLY in one buffer all ranks share (POSIX shared memory,
page-locked by every rank for the part it writes), next to two tables in BATCH ORDER -- offset and count of
every utterance.  When the ranks are done, any process that maps the buffer holds the whole batch.
"""
from __future__ import annotations

import os

import numpy as np


class SharedBatch:
    """[offsets u64 x n | counts u32 x n | pad | pcm int16 x n_samples] in /dev/shm/<name>."""

    def __init__(self, name: str, n_utts: int, n_samples: int, create: bool):
        self.path = os.path.join("/dev/shm", name)
        self.n_utts, self.n_samples = int(n_utts), int(n_samples)
        head = 12 * self.n_utts
        self._pcm_at = (head + 4095) // 4096 * 4096          # page aligned: ranks pin disjoint page ranges
        size = self._pcm_at + 2 * max(self.n_samples, 8)
        if create:
            with open(self.path, "wb") as f:
                f.truncate(size)
        self._mm = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(size,))
        self.offsets = self._mm[:8 * self.n_utts].view(np.uint64)
        self.counts = self._mm[8 * self.n_utts:12 * self.n_utts].view(np.uint32)
        self.pcm = self._mm[self._pcm_at:self._pcm_at + 2 * max(self.n_samples, 8)].view(np.int16)
        self._pinned = None

    def region(self, lo: int, hi: int) -> np.ndarray:
        return self.pcm[int(lo):int(hi)]

    def pin(self, lo: int, hi: int) -> bool:
        """Page-lock samples [lo, hi) for this process (cudaHostRegister); False if that is not possible."""
        import torch
        a = self.pcm[int(lo):int(hi)]
        if a.size == 0:
            return True
        start = a.ctypes.data // 4096 * 4096
        end = (a.ctypes.data + a.nbytes + 4095) // 4096 * 4096
        rc = torch.cuda.cudart().cudaHostRegister(start, end - start, 0)
        ok = int(rc) == 0
        if ok:
            self._pinned = (start, end - start)
        return ok

    def unpin(self) -> None:
        if self._pinned:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self._pinned[0])
            self._pinned = None

    def publish(self, indices, base: int, local_offsets, local_counts) -> None:
        """Rank-local results -> the batch-order tables."""
        idx = np.asarray(indices, dtype=np.int64)
        self.offsets[idx] = np.asarray(local_offsets, dtype=np.uint64) + np.uint64(base)
        self.counts[idx] = np.asarray(local_counts, dtype=np.uint32)

    def utterance(self, u: int) -> np.ndarray:
        o, c = int(self.offsets[u]), int(self.counts[u])
        return self.pcm[o:o + c]

    def close(self, unlink: bool = False) -> None:
        self.unpin()
        self.offsets = self.counts = self.pcm = None
        del self._mm
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def text_costs(texts: list[str], speeds) -> np.ndarray:
    """What an utterance costs BEFORE it is planned: output samples scale with characters / speed (they are
    what crosses PCIe), and an utterance that goes through WSOLA costs its pre-stretch length on top."""
    s = np.clip(np.asarray(speeds, dtype=np.float64), 0.5, 2.0)
    chars = np.array([max(len(t), 1) for t in texts], dtype=np.float64)
    return np.where(s != 1.0, chars / s + chars, chars)


def region_bases(used: list[int]) -> np.ndarray:
    """Start of every rank's region in the shared buffer (8-sample aligned), and the total."""
    b = np.zeros(len(used) + 1, dtype=np.int64)
    for r, u in enumerate(used):
        b[r + 1] = b[r] + (int(u) + 7) // 8 * 8 + 4096
    return b
