"""One process per slab, linked through CUDA IPC (blbm_export_peer / blbm_link_peer) — the deployment
bench.py uses under torchrun.  Runs on a single-GPU box too: both processes then share cuda:0, which still
exercises the cross-process mapping, the direct halo stores and the epoch handshake."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scenario(w, h):
    """(paint list, steps) pairs applied in order by every rank and by the oracle."""
    b = h // 2
    loc = [y * w + x for y in (b - 2, b - 1, b, b + 1) for x in (0, 1, w // 2, w - 2, w - 1)]
    loc += [(h // 3 + dy) * w + (w // 4 + dx) for dy in range(-3, 4) for dx in range(-3, 4)]
    p1 = np.array([[l, 1] for l in loc], np.uint64)
    p2 = np.array([[l, 0] for l in loc[::2]], np.uint64)
    return [(p1, 33), (p2, 20), (None, 1), (None, 14)]


_RESETS = [("custom_speed", 0.07, 12), ("single_cell", 5, 9), ("single_cell", 4, 3), ("custom_speed", 0.1, 1),
           ("single_cell", 3, 21)]


def _worker(rank, world, port, w, h, kernel, lazy, out_dir, resets=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lbm_b200 import LBM, load_library
        from lbm_b200.lbm import slab_rows
        ndev = load_library().blbm_device_count()
        dev = rank % max(ndev, 1)
        rows = slab_rows(h, world)[rank]
        lbm = LBM(1.0 / (3 * 0.02 + 0.5), w, h, device=dev, rows=rows, kernel=kernel, lazy_barriers=lazy)
        blobs = [None] * world
        dist.all_gather_object(blobs, lbm.export_peer())
        if rank > 0:
            lbm.link_peer(0, blobs[rank - 1])
        if rank < world - 1:
            lbm.link_peer(1, blobs[rank + 1])
        dist.barrier()
        for pairs, steps in _scenario(w, h):
            if pairs is not None:
                lbm.draw_points(pairs)
            lbm.iterate(steps)
        if resets:
            # resets are collective: each slab refills its own halo rows, and no neighbour's next launch may store
            # into them before that fill ran (the ranks drift freely here); single_cell 3..5 sit on y/2
            for op, arg, steps in _RESETS:
                getattr(lbm, op)(arg)
                lbm.iterate(steps)
        res = {f"f{b}_{k}": lbm.read_population(k, b) for b in (0, 1) for k in range(9)}
        mx, my, rho = lbm.read_moments()
        res.update(mx=mx, my=my, rho=rho, out=lbm.read_output(), bar=lbm.read_barrier(), cls=lbm.read_cell_class())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
        dist.barrier()  # keep every pool mapped until all neighbours are done with it
        lbm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lazy", [0, 1])
@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("world", [2, 3])
def test_ipc_linked_slabs_match_oracle(world, kernel, lazy, tmp_path):
    from oracle.lbm_oracle import Oracle
    from tests.util import assert_same_bits
    w, h = 160, 45
    mp.spawn(_worker, args=(world, _free_port(), w, h, kernel, lazy, str(tmp_path)), nprocs=world, join=True)
    o = Oracle(1.0 / (3 * 0.02 + 0.5), w, h)
    for pairs, steps in _scenario(w, h):
        if pairs is not None:
            o.draw_points(pairs.astype(np.uint32))
        o.iterate(steps)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    cat = lambda name: np.concatenate([p[name] for p in parts], axis=0)  # noqa: E731
    for b in (0, 1):
        for k in range(9):
            assert_same_bits(cat(f"f{b}_{k}"), o.population(b, k), f"population[{b}][{k}]")
    omx, omy, orho = o.moments()
    assert_same_bits(cat("mx"), omx, "mx")
    assert_same_bits(cat("my"), omy, "my")
    assert_same_bits(cat("rho"), orho, "rho")
    assert_same_bits(cat("out"), o.output(), "curl output")
    assert_same_bits(cat("bar"), o.barrier(), "barrier")
    assert_same_bits(cat("cls"), o.cell_class(), "class")


@pytest.mark.parametrize("world", [2, 3])
def test_ipc_linked_slabs_resets_are_ordered_against_neighbour_pushes(world, tmp_path):
    """custom_speed / single_cell on linked slabs of independent processes, each followed by iterate: the refill of
    a slab's halo rows is published through the epoch handshake, so a neighbour that is already past its own reset
    cannot have its first halo stores overwritten by our fill."""
    from oracle.lbm_oracle import Oracle
    from tests.util import assert_same_bits
    w, h = 160, 46
    mp.spawn(_worker, args=(world, _free_port(), w, h, 2, 1, str(tmp_path), True), nprocs=world, join=True)
    o = Oracle(1.0 / (3 * 0.02 + 0.5), w, h)
    for pairs, steps in _scenario(w, h):
        if pairs is not None:
            o.draw_points(pairs.astype(np.uint32))
        o.iterate(steps)
    for op, arg, steps in _RESETS:
        getattr(o, op)(arg)
        o.iterate(steps)
    parts = [np.load(os.path.join(str(tmp_path), f"rank{r}.npz")) for r in range(world)]
    cat = lambda name: np.concatenate([p[name] for p in parts], axis=0)  # noqa: E731
    for b in (0, 1):
        for k in range(9):
            assert_same_bits(cat(f"f{b}_{k}"), o.population(b, k), f"population[{b}][{k}]")
    omx, omy, orho = o.moments()
    assert_same_bits(cat("mx"), omx, "mx")
    assert_same_bits(cat("rho"), orho, "rho")
    assert_same_bits(cat("out"), o.output(), "curl output")
