"""world_size-2 gloo test of the sharding + single all-gather used by the multi-GPU path."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from gimmvfi_b200.parallel import shard_range

    for n in (0, 1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_gather_frames_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import sys, torch
        sys.path.insert(0, %r)
        from gimmvfi_b200.parallel import init_from_env, shard_range, gather_frames
        rank, local, world = init_from_env("gloo")
        for n_pairs in (4, 5):
            mine = shard_range(n_pairs, rank, world)
            frames = torch.stack([torch.full((3, 4, 6), float(i)) for i in mine]) if len(mine) else torch.zeros(0, 3, 4, 6)
            allf = gather_frames(frames, n_pairs)
            assert allf.shape == (n_pairs, 3, 4, 6), allf.shape
            assert [int(allf[i, 0, 0, 0]) for i in range(n_pairs)] == list(range(n_pairs))
        print("rank", rank, "ok")
    """ % ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
