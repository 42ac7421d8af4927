"""world_size-2 gloo test (CPU) of the N>1 host logic: VBlocks are sharded round-robin by vblock_i with no data-path
collective; the only exchange is the final section-list gather."""
import os, sys
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import genozip_b200
    L = genozip_b200.load()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_vb = 11
    mine = [v for v in range(1, n_vb + 1) if L.gzb_vb_device(v, world) == rank]
    # each rank "compresses" its VBlocks: here the section list is (vblock_i, n_sections, total_len) per VB
    sec = torch.zeros((n_vb, 3), dtype=torch.int64)
    for v in mine:
        sec[v - 1] = torch.tensor([v, 9, 1000 + 7 * v])
    out = [torch.zeros_like(sec) for _ in range(world)]
    dist.all_gather(out, sec)                      # final section-list gather
    full = torch.stack(out).sum(0)
    q.put((rank, mine, full.tolist()))
    dist.destroy_process_group()


def test_round_robin_and_section_list_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(30) for p in ps]
    owned = sorted(v for _, mine, _ in res for v in mine)
    assert owned == list(range(1, 12)), "every VBlock owned exactly once"
    for rank, mine, full in res:
        assert all((v - 1) % world == rank for v in mine)
        assert [row[0] for row in full] == list(range(1, 12)), "gathered section list is complete on every rank"
        assert full == res[0][2]
