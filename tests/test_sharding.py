"""CPU: the multi-GPU path is frame sharding with no data-path collective.  A world_size-2 (and 3)
gloo job checks the partition and that sharded graph construction equals a single-process pass."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_block_partition(built):
    from toucan_b200 import render, workloads

    for frames in (1, 7, 10, 1000):
        for world in (1, 2, 3, 4, 8):
            blocks = [workloads.frame_block(r, world, frames) for r in range(world)]
            assert blocks == [render.frame_block(0, frames, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == frames
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(hi - lo for lo, hi in blocks) == -(-frames // world)


def test_stack_layers_match_oracle_graph():
    # the bench's closed-form layer list equals what the oracle's ImageGraph restatement builds
    import numpy as np

    from oracle import graph as G
    from toucan_b200 import workloads

    cfg = dict(tracks=2, frames=150)
    tl = G.Timeline(workloads.stack_timeline(**cfg))
    g = G.ImageGraph(tl, lambda p: np.zeros((4, 4, 4), np.float32))
    for f in range(150):
        tree = "\n".join(G.describe_tree(g.exec(G.RT(float(f), 24.0))))
        layers = workloads.stack_layers(f, **cfg)
        assert tree.count("toucan:Dissolve") == sum(la[1] is not None for la in layers), f
        for (ta, ka), b, ra, rb, v in layers:
            assert f"Read: still{ka}.exr" in tree
            if b is not None:
                assert f"value={G._fmt_any(float(v))}" in tree, (f, v)


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_graph_construction(built, world):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world + os.getpid() % 200), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"SHARDING_OK {world}" in r.stdout
