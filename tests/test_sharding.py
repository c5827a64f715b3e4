"""Multi-GPU host logic on CPU: utterance sharding (no data-path collective) with a world_size-2 gloo group."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_indices_partition(H):
    sh = H.pkg.sharding
    rng = np.random.default_rng(0)
    costs = rng.uniform(1, 100, 1000)
    for world in (1, 2, 3, 8):
        shards = sh.shard_indices(costs, world)
        cat = np.concatenate(shards)
        assert sorted(cat.tolist()) == list(range(1000))
        loads = [costs[s].sum() for s in shards]
        assert max(loads) - min(loads) <= costs.max() + 1e-9     # LPT balance bound
        inv = sh.gather_order(shards)
        assert np.array_equal(cat[inv], np.arange(1000))


def _worker(rank: int, world: int, port: int, ret):
    import harness as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fr = H.front.Front(H.small_db(), H.shipped_config(), H.NORM_CSV)
        texts = H.corpus.batch(32, seed=3)
        speeds = np.where(np.arange(32) % 4 == 0, 1.5, 1.0).astype(np.float32)
        plan = fr.plan(texts, speeds)
        pre, out, _ = fr.bounds(plan)
        sh = H.pkg.sharding
        shards = sh.shard_indices(sh.utterance_costs(pre, speeds), world)
        mine = plan.select(shards[rank])
        # stand-in for the GPU stage: the per-utterance upper bounds of this rank's shard
        local = torch.from_numpy(fr.bounds(mine)[1].astype(np.int64))
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([len(local)]))
        bufs = [torch.zeros(int(s), dtype=torch.int64) for s in sizes]
        # host-side gather of per-rank results (gloo all_gather needs equal sizes: pad)
        m = int(max(int(s) for s in sizes))
        padded = torch.zeros(m, dtype=torch.int64)
        padded[:len(local)] = local
        got = [torch.zeros(m, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, padded)
        cat = torch.cat([g[:int(s)] for g, s in zip(got, sizes)]).numpy()
        merged = cat[sh.gather_order(shards)]
        # time = max over ranks, work = sum over ranks (what bench.py reduces)
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w = torch.tensor([float(local.sum())])
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        ok = np.array_equal(merged, out.astype(np.int64)) and float(t) == world and float(w) == float(out.sum())
        if rank == 0:
            ret.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True


def _gather_worker(rank: int, world: int, port: int, name: str, ret):
    """The host-side gather of bench.py's strong-scaling leg with a stand-in for the device: every rank writes
    its shard's (fake) PCM into its own region of the shared buffer and publishes offsets / counts in batch order."""
    import harness as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hg, sh = H.pkg.hostgather, H.pkg.sharding
        texts = H.corpus.batch(64, seed=9, target_chars=60)
        speeds = H.corpus.mixed_speeds(64, seed=2)
        shards = sh.shard_indices(hg.text_costs(texts, speeds), world)
        mine = shards[rank]
        counts = np.array([100 + 3 * len(texts[i]) for i in mine], dtype=np.int64)       # "samples" of my utterances
        slots = (counts + 7) // 8 * 8 + 8
        used = int(slots.sum())
        all_used = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_used, torch.tensor([used]))
        bases = hg.region_bases([int(u) for u in all_used])
        if rank == 0:
            sb = hg.SharedBatch(name, len(texts), int(bases[-1]), create=True)
        dist.barrier()
        if rank != 0:
            sb = hg.SharedBatch(name, len(texts), int(bases[-1]), create=False)
        region = sb.region(bases[rank], bases[rank] + used)
        off = np.concatenate([[0], np.cumsum(slots)[:-1]])
        for k, i in enumerate(mine):
            region[off[k]:off[k] + counts[k]] = (np.arange(counts[k]) + 7 * i) % 30000
        sb.publish(mine, int(bases[rank]), off, counts)
        dist.barrier()
        ok = True
        if rank == 0:
            for i in range(len(texts)):
                want = (np.arange(100 + 3 * len(texts[i])) + 7 * i) % 30000
                ok = ok and np.array_equal(sb.utterance(i), want.astype(np.int16))
            ret.put(bool(ok))
        dist.barrier()
        sb.close(unlink=rank == 0)
    finally:
        dist.destroy_process_group()


def test_two_rank_host_gather_through_shared_memory():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    name = f"ctts_b200_test_{os.getpid()}"
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, name, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True
    assert not os.path.exists(os.path.join("/dev/shm", name))
