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


# ------------------------------------------------------------------------------------------------
# Keypoint long-video windows sharded over ranks: one all-reduce of (value, count) per step
# ------------------------------------------------------------------------------------------------
def _window_setup():
    import ref_ops
    from mofa_video_b200.pipeline import svdxt_pipeline_ctrlnet_loop as kpl
    from oracle import fixtures
    from test_keypoint_cpu import _loop_setup, to_cl
    cfg = dict(fixtures.TINY_CONFIG)
    H = W = 16
    T, F_frames, stride = cfg["num_frames"], 7, 2
    s = _loop_setup(cfg, H, W, F_frames)
    views = kpl.unique_views(kpl.window_views(F_frames, T, stride))
    states = []
    for (ts, te), mult in views:
        fl = s["flow"][0, (ts - 1):(te - 1)].half().contiguous()
        lm = torch.cat([s["ldmk"][0, 0:1], s["ldmk"][0, ts:te]])
        s["f_net"].adapter_cond_branch_ldmk(to_cl(s["cond"][:1]), fl, to_cl(lm), 8 * H, 8 * W)
        states.append(((ts, te), mult, (s["f_net"].warped, s["f_net"].ldmk)))
    lat = s["lat0"][0].half().reshape(F_frames, 4, H * W).contiguous()
    il = s["il"][:, 0].half().reshape(2, 4, H * W).contiguous()
    args = (ref_ops, s["u_net"], s["f_net"], states, lat, il, s["sch"]._sigmas_host, s["sch"]._timesteps_host, H, W, T,
            1.0, 3.0, 0.9)
    return kpl, args, len(views)


def _window_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kpl, args, n_views = _window_setup()
    states = [st if k % world == rank else None for k, st in enumerate(args[3])]   # a rank only prepares its own windows
    out = kpl.denoise_windowed(*args[:3], states, *args[4:], shard=(world, rank))
    torch.save({"lat": out, "n_views": n_views}, os.path.join(out_dir, f"w{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_keypoint_windows_sharded_over_two_ranks(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    kpl, args, n_views = _window_setup()
    assert n_views >= 3 and kpl.views_for_rank(n_views, 2, 1) == list(range(1, n_views, 2))
    single = kpl.denoise_windowed(*args)
    port = _free_port()
    mp.spawn(_window_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w0, w1 = torch.load(tmp_path / "w0.pt"), torch.load(tmp_path / "w1.pt")
    assert torch.equal(w0["lat"], w1["lat"])                       # every rank ends with the same latents
    err = (w0["lat"].float() - single.float()).abs().max() / single.float().abs().max()
    assert err < 2e-3, err                                          # = the single-process loop (fp32 sum order only)
