"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path — rendezvous, broadcast of the
128-byte communicator id, and the row / problem / sample partitions (they must tile the index space
exactly, whatever the rank count).  The data path itself (NCCL all-gather inside the CUDA library) is
covered on the GPU box by tests/dist_check.py."""
import importlib.util
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_dist():
    spec = importlib.util.spec_from_file_location("osb_dist", os.path.join(ROOT, "optimization-solvers_b200", "dist.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    d = _load_dist()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    payload = bytes(range(128)) if rank == 0 else None
    got = d.broadcast_bytes(payload)
    rows = d.shard_rows(16384, rank, world)
    probs = d.shard_problems(262144 + 3, rank, world)
    samp = d.shard_samples(1 << 20, rank, world)
    allr = [None] * world
    dist.all_gather_object(allr, (rows, probs, samp, got == bytes(range(128)), d.env_rank_world()))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put(allr)


def test_partitions_and_id_broadcast_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    allr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[3] for r in allr)
    assert [r[4][0] for r in allr] == [0, 1]
    # partitions tile the index space
    assert allr[0][0] == (0, 8192) and allr[1][0] == (8192, 8192)
    p = [r[1] for r in allr]
    assert p[0][0] == 0 and p[0][0] + p[0][1] == p[1][0] and p[1][0] + p[1][1] == 262144 + 3
    assert allr[0][2] == (0, 1 << 19) and allr[1][2] == (1 << 19, 1 << 19)


def test_partition_rules():
    d = _load_dist()
    for world in (1, 2, 4, 8):
        spans = [d.shard_rows(16384, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(s[1] for s in spans) == 16384
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        pr = [d.shard_problems(1000, r, world) for r in range(world)]
        assert sum(c for _, c in pr) == 1000 and all(pr[i][0] + pr[i][1] == pr[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        d.shard_rows(100, 0, 8)
    with pytest.raises(ValueError):
        d.shard_samples(10, 0, 4)


def test_index_range_sharding_rules():
    """Index-range sharding of GD / PGD / SPG: equal contiguous ranges, whole functor blocks."""
    d = _load_dist()
    for world in (1, 2, 4, 8):
        spans = [d.shard_indices(1 << 28, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == 1 << 28
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert len({c for _, c in spans}) == 1 and spans[0][1] % 2 == 0
    with pytest.raises(ValueError):
        d.shard_indices(1002, 0, 4)  # 1002 / 4 is not a whole number of 2-blocks


def test_rank_ordered_combine_matches_the_device_rule_world2(tmp_path):
    """The cross-rank combine of the sharded reductions is 'fold the per-rank values in rank order' (sum, max, min):
    two gloo ranks all-gather their partials and fold them; both obtain the same bits, equal to the sequential fold."""
    import subprocess
    import sys
    code = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(7)
parts = rng.standard_normal((world, 3)) * [1.0, 1e8, 1e-8]
mine = torch.from_numpy(parts[rank].copy())
allp = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
dist.all_gather(allp, mine)
def fold(ps):
    s, mx, mn = 0.0, -np.inf, np.inf
    for p in ps:
        s = s + float(p[0]); mx = max(mx, float(p[1])); mn = min(mn, float(p[2]))
    return s, mx, mn
got = fold(allp)
want = fold([torch.from_numpy(parts[r]) for r in range(world)])
assert got == want, (got, want)
res = [None] * world
dist.all_gather_object(res, got)
assert all(r == res[0] for r in res)
dist.destroy_process_group()
'''
    script = tmp_path / "combine_worker.py"
    script.write_text(code)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]


def test_packed_symmetric_layouts_tile_the_storage_exactly():
    """Host logic of the packed storage, single GPU and sharded by tile pairs: every tile has
    exactly one owner, the tiles of a rank are disjoint, in bounds and leave no gap, and the ranks' loads are equal."""
    import importlib
    sys.path.insert(0, ROOT)
    osb = importlib.import_module("optimization-solvers_b200")
    for n in (16, 128, 2048, 16384):
        T = (n + 7) // 8
        for world in (1, 2, 4, 8):
            if world > 1 and n // 16 < world:
                continue
            spans = {r: [] for r in range(world)}
            totals = {}
            for tile in range(T):
                for r in range(world):
                    owner, off, ldp, tot = osb.sym_layout(n, world, r, tile)
                    totals[r] = tot
                assert 0 <= owner < world and ldp >= 8 * (tile + 1) and ldp % 16 == 0
                spans[owner].append((off, off + 8 * ldp))
            for r in range(world):
                sp = sorted(spans[r])
                # (the single-GPU array carries one spare zero tile after the last one; the sharded arrays are exact)
                assert sp[0][0] == 0 and (sp[-1][1] == totals[r] if world > 1 else sp[-1][1] <= totals[r])
                assert all(sp[i][1] == sp[i + 1][0] for i in range(len(sp) - 1))  # no overlap, no gap
            if world > 1:
                assert len(set(totals.values())) == 1 or max(totals.values()) - min(totals.values()) <= 8 * (8 * T + 16)
    with pytest.raises(osb.ErrorInputParams):
        osb.sym_layout(24, 2, 0, 0)  # odd tile count cannot be sharded by pairs


def test_snake_dealing_of_local_tiles_covers_and_balances():
    """Restatement of the tile schedule of qn_lazy_sym_kernel<.., SHARDED> (csrc/qn_kernels.cu): a rank's tiles, in
    decreasing length, are dealt to its CTAs in snake order.  Every local tile is visited exactly once, and the
    heaviest CTA carries at most max(longest tile, mean + one tile) columns — the quantisation figures quoted in
    DESIGN.md section 6 (0.865 at 8 GPUs, near 1 at 2 GPUs) follow from it."""
    n, G = 16384, 148
    T = n // 8
    for world in (2, 4, 8):
        for rank in range(world):
            half = T // 2
            nlp = (half - rank + world - 1) // world
            nunits = 2 * nlp
            grid = min(nunits, G)
            seen, load = [], [0] * grid
            for b in range(grid):
                k = 0
                while k * grid < nunits:
                    pos = grid - 1 - b if (k & 1) else b
                    uq = k * grid + pos
                    k += 1
                    if uq >= nunits:
                        continue
                    pairi = rank + uq * world if uq < nlp else rank + (2 * nlp - 1 - uq) * world
                    tile = T - 1 - pairi if uq < nlp else pairi
                    seen.append(tile)
                    load[b] += 8 * (tile + 1)
            owned = sorted(t for t in range(T) if min(t, T - 1 - t) % world == rank)
            assert sorted(seen) == owned
            mean = sum(load) / grid
            assert max(load) <= max(n, mean + n / 2), (world, rank, max(load), mean)
            if world == 2:
                assert max(load) / mean < 1.03
            if world == 8:
                assert abs(mean / max(load) - 0.865) < 0.01
