"""-m gpu (one GPU, one process): the peer-store frame gather's kernels and protocol with world = 1 -- the decoder tail
writes into the gather slot (here local memory; on N > 1 the same pointer arithmetic lands in rank 0's HBM over NVLink,
covered by tools/peer_gather_check.py under torchrun), flags publish / collect / release across clips, the u8x4 tail
kernel bit-exact against the scalar one, and the time-out path of mofa_peer_wait."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tail_kernel_u8x4_equals_scalar_and_statement():
    from mofa_video_b200 import lib
    import ref_ops
    T, H, W = 5, 24, 40
    g = torch.Generator().manual_seed(0)
    y = (torch.randn(T * H * W, 3, generator=g) * 1.2).half().cuda()
    w = (torch.randn(3, 3, 3, generator=g) * 0.4).cuda()
    b = (torch.randn(3, generator=g) * 0.1).cuda()
    u4 = torch.zeros(T, H, W, 3, dtype=torch.uint8, device="cuda")
    lib.vae_time_conv_out(y, w, b, None, u4, T, H * W)                    # u8-only, HW % 4 == 0 -> x4 kernel
    f32 = torch.empty(T, 3, H, W, device="cuda")
    u1 = torch.zeros_like(u4)
    lib.vae_time_conv_out(y, w, b, f32, u1, T, H * W)                     # scalar kernel (also writes fp32)
    assert torch.equal(u4, u1)
    rf, ru = torch.empty(T, 3, H, W), torch.zeros(T, H, W, 3, dtype=torch.uint8)
    ref_ops.vae_time_conv_out(y.cpu(), w.cpu(), b.cpu(), rf, ru, T, H * W)
    assert (u4.cpu().int() - ru.int()).abs().max().item() <= 1


def test_peer_gather_protocol_single_rank():
    from mofa_video_b200 import lib
    from mofa_video_b200.parallel import PeerFrameGather
    T, H, W = 3, 16, 24
    gth = PeerFrameGather((T, H, W, 3))
    g = torch.Generator().manual_seed(1)
    w = (torch.randn(3, 3, 3, generator=g) * 0.4).cuda()
    b = torch.zeros(3).cuda()
    for clip in range(3):
        y = (torch.randn(T * H * W, 3, generator=g)).half().cuda()
        want = torch.zeros(T, H, W, 3, dtype=torch.uint8, device="cuda")
        lib.vae_time_conv_out(y, w, b, None, want, T, H * W)
        slot = gth.begin()                                               # waits for release of clip - 1
        lib.vae_time_conv_out(y, w, b, None, slot, T, H * W)
        gth.publish()
        allf = gth.collect()
        assert allf.shape == (1, T, H, W, 3)
        got = allf.clone()
        gth.release()
        torch.cuda.synchronize()
        gth.check()
        assert torch.equal(got[0], want)
        assert int(gth.ready[0].item()) == clip + 1 and int(gth.consumed.item()) == clip + 1
    # a flag nobody sets: the wait gives up after its time-out instead of hanging the GPU
    flags = torch.zeros(2, dtype=torch.int32, device="cuda")
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    lib.peer_signal(flags[0:1], 5)
    lib.peer_wait(flags, 5, timeout_s=0.2, timed_out=bad)
    torch.cuda.synchronize()
    assert int(bad.item()) == 2                                           # 1 + index of the silent flag
