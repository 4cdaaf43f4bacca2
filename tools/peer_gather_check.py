"""torchrun --nproc-per-node N tools/peer_gather_check.py : the peer-store frame gather across real GPUs.
Every rank runs the decoder tail kernel into ITS slot of rank 0's buffer (NVLink stores issued by the kernel itself) for a
few clips; rank 0 compares every slot with what that rank computed locally (sent over NCCL as the checker) and times the
gather of 576x1024x25 frames both ways."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mofa_video_b200 import lib, parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib.load()
    T, H, W = 25, 576, 1024
    gth = parallel.PeerFrameGather((T, H, W, 3))
    g = torch.Generator().manual_seed(7 + rank)
    w = (torch.randn(3, 3, 3, generator=g) * 0.4).to(dev)
    b = torch.zeros(3, device=dev)
    ok = True
    for clip in range(3):
        want = torch.empty(T, H, W, 3, dtype=torch.uint8, device=dev)
        slot = gth.begin()
        for c0 in range(0, T, 8):                                   # chunks of 8 frames, like decode_chunk_size = 8
            n = min(8, T - c0)
            y = torch.randn(n * H * W, 3, generator=g).half().to(dev)
            lib.vae_time_conv_out(y, w, b, None, want[c0:c0 + n], n, H * W)
            lib.vae_time_conv_out(y, w, b, None, slot[c0:c0 + n], n, H * W)
        gth.publish()
        ref = parallel.gather_frames(want, dst=0)                   # NCCL: the checker
        if rank == 0:
            allf = gth.collect()
            for r in range(world):
                same = torch.equal(allf[r], ref[r])
                ok &= same
                print(f"clip {clip} slot {r}: {'identical' if same else 'MISMATCH'}", flush=True)
            gth.release()
        torch.cuda.synchronize()
        gth.check()
    # timing: tail kernel writing locally + NCCL gather  vs  tail kernel writing straight into rank 0's buffer
    y = torch.randn(8 * H * W, 3, generator=g).half().to(dev)
    local_out = torch.empty(T, H, W, 3, dtype=torch.uint8, device=dev)

    def via_nccl():
        for c0 in range(0, T, 8):
            n = min(8, T - c0)
            lib.vae_time_conv_out(y[: n * H * W], w, b, None, local_out[c0:c0 + n], n, H * W)
        parallel.gather_frames(local_out, dst=0)

    def via_peer():
        slot = gth.begin()
        for c0 in range(0, T, 8):
            n = min(8, T - c0)
            lib.vae_time_conv_out(y[: n * H * W], w, b, None, slot[c0:c0 + n], n, H * W)
        gth.publish()
        if rank == 0:
            gth.collect()
            gth.release()

    for name, fn in (("tail + NCCL gather", via_nccl), ("tail with peer stores", via_peer)):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{name}: {ms.item():.3f} ms per clip (max over {world} ranks; {T * H * W * 3 / 1e6:.1f} MB per rank)",
                  flush=True)
    gth.check()
    if rank == 0:
        print("PEER_GATHER_OK" if ok else "PEER_GATHER_FAILED", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
