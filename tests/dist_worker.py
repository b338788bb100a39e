"""Worker of tests/test_sharding.py: one rank of a world_size-N gloo job on the CPU.

Exercises the N > 1 logic of bench.py / the frame scheduler without a GPU: every rank takes
its contiguous block of frames, builds the node tree of every frame of its block with the C++
host, and the ranks gather (frame, digest) pairs; rank 0 checks the blocks tile the timeline
exactly once, in order, and that the gathered trees equal a single-process pass."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch.distributed as dist

    from toucan_b200 import render, workloads

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    frames = 200
    doc = workloads.stack_timeline(tracks=3, frames=frames)
    tl = render.Timeline(json_doc=doc, directory="/synthetic")
    for t in range(3):
        for k in range(2):
            tl.set_media(tl.media_path(workloads.still_url(t, k)), np.zeros((8, 8, 4), np.float32))
    tl.prepare()
    host = render.Host()
    lo, hi = workloads.frame_block(rank, world, frames)
    assert (lo, hi) == render.frame_block(0, frames, rank, world), "python and C++ partitions differ"
    mine = [(f, hashlib.sha256(tl.describe(host, float(f)).encode()).hexdigest()) for f in range(lo, hi)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        flat = [x for part in gathered for x in part]
        assert [f for f, _ in flat] == list(range(frames)), "blocks must tile the timeline once, in order"
        single = [hashlib.sha256(tl.describe(host, float(f)).encode()).hexdigest() for f in range(frames)]
        assert [d for _, d in flat] == single
        # every rank's bench steps stay inside its block and hold the same 50:50 mix of plain and transition frames; its
        # end-to-end window is consecutive frames of the block, half of them inside a transition
        sys.path.insert(0, ROOT)
        import bench

        cfg = dict(workloads.C5)
        for r in range(world):
            fr = bench.c5_frames(r, world, cfg, 24)
            lo_r, hi_r = workloads.frame_block(r, world, cfg["frames"])
            assert all(lo_r <= f < hi_r for f in fr)
            inside = [workloads.stack_layers(f, **cfg)[0][1] is not None for f in fr]
            assert inside == [False, True] * 12, inside
            first = bench.c5_e2e_window(r, world, cfg, 20)
            assert lo_r <= first and first + 20 <= hi_r
            window = [workloads.stack_layers(f, **cfg)[0][1] is not None for f in range(first, first + 20)]
            assert window == [False] * 10 + [True] * 10, window
        print("SHARDING_OK", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
