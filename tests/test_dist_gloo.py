"""The N>1 path on CPU: world_size-2 gloo run of the step's only cross-rank exchange
(all-reduce of the patch gradient + all-gather of the per-sample results) and of the
EOT-shard slicing: the sharded result must equal the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dorpatch_b200.attack import exchange_shards
    B, S, HW = 3, 8, 5
    rng = np.random.RandomState(0)
    g_all = rng.rand(world, B, 3, HW, HW).astype(np.float32)          # per-rank partial gradient
    loss = rng.rand(B, S).astype(np.float32)
    preds = rng.randint(0, 1000, (B, S)).astype(np.int32)
    s_loc = S // world
    sl = slice(rank * s_loc, (rank + 1) * s_loc)
    G = torch.from_numpy(g_all[rank].copy())
    la, pa = exchange_shards(dist, G, loss[:, sl].copy(), preds[:, sl].copy())
    ok = np.allclose(G.numpy(), g_all.sum(0), rtol=1e-6) and np.array_equal(la, loss) and np.array_equal(pa, preds)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_shards_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _scan_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dorpatch_b200.attack import scan_failures
    ok = True
    for n, y, targeted in ((2520, 3, False), (145, 1, True), (1, 0, False)):   # even split, ragged last shard, empty shard
        rects = np.zeros((n, 4, 4), np.int16)
        rects[:, 0, 0] = np.arange(n)                                     # tag every universe entry with its index
        labels = np.random.RandomState(n).randint(0, 5, n)
        seen = []

        def predict(r):
            seen.append(r[:, 0, 0].copy())
            return labels[r[:, 0, 0]]

        want = scan_failures(lambda r: labels[r[:, 0, 0]], rects, y, targeted)
        got = scan_failures(predict, rects, y, targeted, dist, "cpu")
        per = -(-n // world)
        mine = np.arange(min(rank * per, n), min((rank + 1) * per, n))
        ok = ok and got == want and (np.array_equal(np.concatenate(seen), mine) if len(mine) else not seen)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_universe_scan_is_sharded_by_mask_index_world2():
    """collect_failure under torch.distributed: every rank scans only its contiguous shard of the mask
    universe and all ranks end with the single-process failure list."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scan_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _consistency_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dorpatch_b200.attack import _assert_same_on_all_ranks
    idx = np.random.RandomState(3).randint(0, 2520, (4, 16))
    _assert_same_on_all_ranks(dist, idx, "cpu", "identical indices")          # must pass
    bad = idx.copy()
    if rank == 1:
        bad[2, 5], bad[2, 6] = bad[2, 6], bad[2, 5]                             # same multiset, different order
    try:
        _assert_same_on_all_ranks(dist, bad, "cpu", "diverged indices")
        raised = False
    except RuntimeError:
        raised = True
    q.put((rank, raised))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_divergence_of_sample_indices_is_detected_world2():
    """EOT sharding assumes every rank drew the same sample indices; a rank with another RNG state must be caught
    (on every rank, so that no rank is left waiting in a collective)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_consistency_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
