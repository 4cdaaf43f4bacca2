"""-m "not gpu": the N>1 path (clip sharding + final uint8 frame gather) with world_size 2 over gloo on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mofa_video_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.clips_for_rank(5, world, rank)
    # each rank "renders" its first clip: frames filled with the clip index
    frames = torch.full((3, 4, 6, 3), mine[0], dtype=torch.uint8)
    got = parallel.gather_frames(frames, dst=0)
    if rank == 0:
        ok = got is not None and len(got) == world and all(int(g.max()) == r and int(g.min()) == r
                                                             for r, g in enumerate(got))
        torch.save({"ok": ok, "mine": mine}, os.path.join(out_dir, "r0.pt"))
    else:
        torch.save({"ok": got is None, "mine": mine}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_clip_sharding_and_gather_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert r0["ok"] and r1["ok"]
    assert r0["mine"] == [0, 2, 4] and r1["mine"] == [1, 3]


def test_single_process_gather_is_identity():
    f = torch.zeros(2, 2, 2, 3, dtype=torch.uint8)
    assert parallel.gather_frames(f)[0] is f
