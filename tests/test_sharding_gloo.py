"""N > 1 host-side logic on CPU: world_size-2 gloo run of the band sharding + gather + assembly
(raytracer_rs_b200/multi_gpu.py). The pixels come from the CPU oracle here; the GPU path is covered by -m gpu tests
and bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest

from raytracer_rs_b200.multi_gpu import assemble_frame, band_partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_partition_covers_every_row_once():
    for h, world, band in [(1080, 8, 8), (1080, 2, 8), (768, 4, 8), (37, 3, 5), (2160, 8, 8)]:
        parts = band_partition(h, world, band)
        allrows = np.sort(np.concatenate(parts))
        assert np.array_equal(allrows, np.arange(h))
        for r, rows in enumerate(parts):
            assert ((rows // band) % world == r).all()
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= band


def test_interleaving_balances_the_empty_bottom_of_the_frame():
    """Q1 leaves the lower ~38 % of a 16:9 frame empty; contiguous bands would idle half the GPUs, interleaved bands
    give every rank the same share of rows from the top (hit-heavy) 60 %."""
    parts = band_partition(1080, 8, 8)
    top = [int((p < 648).sum()) for p in parts]
    assert max(top) - min(top) <= 8


def _worker(rank, world, port, w, h, result_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import raytracer_rs_b200 as rt
    from oracle_lib import JITTER_FIXED, Oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    scene = rt.load_scene(os.path.join(ROOT, "data", "ico2.dae"))
    parts = band_partition(h, world, 8)
    mine = parts[rank]
    orc = Oracle(scene, w, h)
    orc.configure(recursions=0, jitter=JITTER_FIXED)
    for b0 in range(0, len(mine), 8):  # one band at a time, exactly the rows this rank owns
        band = mine[b0:b0 + 8]
        orc.trace_rows(int(band[0]), len(band), 1, threads=1)
    ldr = orc.get_tonemapped_pixels().reshape(h, w)
    max_rows = max(len(p) for p in parts)
    compact = torch.zeros(max_rows * w, dtype=torch.int32)
    compact[: len(mine) * w] = torch.from_numpy(ldr[mine].reshape(-1).view(np.int32))
    recv = [torch.zeros_like(compact) for _ in range(world)] if rank == 0 else None
    dist.gather(compact, recv, dst=0)
    counts = torch.tensor([orc.counters()["rays"]["primary"], orc.counters()["rays"]["shadow"]], dtype=torch.int64)
    dist.all_reduce(counts)
    if rank == 0:
        frame = assemble_frame([r.numpy().view(np.uint32) for r in recv], parts, w, h)
        full = Oracle(scene, w, h)
        full.configure(recursions=0, jitter=JITTER_FIXED)
        full.trace_rows(0, h, 1, threads=1)
        ok = bool(np.array_equal(frame, full.get_tonemapped_pixels()))
        c = full.counters()
        ok_counts = counts.tolist() == [c["rays"]["primary"], c["rays"]["shadow"]]
        with open(result_path, "w") as f:
            f.write(f"{ok} {ok_counts}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_render_gathers_to_rank0(tmp_path, world):
    import torch.multiprocessing as mp

    port = 29500 + (os.getpid() % 2000) + world
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(world, port, 96, 70, str(result)), nprocs=world, join=True)
    assert result.read_text() == "True True"
