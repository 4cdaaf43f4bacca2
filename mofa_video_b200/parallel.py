"""Multi-GPU plumbing for the one place the path shards: independent clips on the batch axis.
One process per GPU (torchrun), weights replicated, clip k -> rank k mod W, no communication during a clip; the only
exchange is the gather of the finished uint8 frames on rank 0.

Two implementations of that gather:
  * PeerFrameGather (CUDA, NVLink / NVSwitch) -- SURVEY.md §8f-3, "frame epilogue fused with the collective": the
    gather buffer [W, T, H, W, 3] lives in rank 0's HBM and is IPC-mapped into every rank; the VAE decoder's tail kernel
    (time_conv_out + clamp + uint8, csrc/elementwise.cu) on rank r stores its frames straight into slot r as they are
    produced, chunk by chunk, overlapping the next chunk's decode.  There is no staging tensor, no NCCL call and no
    extra pass: the bytes cross NVLink exactly once, written by the kernel that computes them.  Two flag words per rank
    (release / acquire at system scope) order producer and consumer.
  * gather_frames -- one `dist.gather` (NCCL on device tensors, gloo on CPU tensors): the portable form, used by the
    CPU tests and where peer access is unavailable.
"""
import torch
import torch.distributed as dist


def clips_for_rank(n_clips, world, rank):
    """Round-robin assignment of clip indices (clip k -> rank k mod world)."""
    return list(range(rank, n_clips, world))


def gather_frames(frames_u8, dst=0):
    """frames_u8: uint8 [T, H, W, 3] of this rank's clip.  Returns the list of all ranks' clips on `dst`
    (None elsewhere).  NCCL on device tensors, gloo on CPU tensors; a single collective, no reduction."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [frames_u8]
    world, rank = dist.get_world_size(), dist.get_rank()
    bufs = [torch.empty_like(frames_u8) for _ in range(world)] if rank == dst else None
    dist.gather(frames_u8.contiguous(), bufs, dst=dst)
    return bufs


class PeerFrameGather:
    """Rank-`dst`-resident gather buffer written by every rank's decoder epilogue over NVLink.

        gather = PeerFrameGather((T, H, W, 3))      # collective: every rank, once
        pipe.frame_sink = gather                    # the pipeline's decode writes into gather.begin()
        out = pipe(..., output_type="uint8_pt")     # frames[0] IS this rank's slot (peer memory on ranks != dst)
        gather.publish()                            # this rank's frames are complete (stream-ordered flag)
        if rank == dst:
            all_u8 = gather.collect()               # [W, T, H, W, 3] on dst, valid on the current stream
            host.copy_(all_u8, non_blocking=True)
            gather.release()                        # lets the ranks overwrite their slots with the next clip

    Flags (int32, in the same allocation, on dst): ready[r] = epoch of the last clip rank r finished writing;
    consumed = epoch of the last clip dst has finished reading.  A writer waits for consumed >= epoch - 1 right before its
    first store of a clip (after the whole denoise loop, so the wait never costs anything in practice)."""

    FLAG_BYTES = 4096

    def __init__(self, frame_shape, dst=0, ops=None, timeout_s=60.0):
        from mofa_video_b200 import lib as _lib
        self.ops = ops if ops is not None else _lib
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst, self.timeout_s = dst, timeout_s
        self.frame_shape = tuple(frame_shape)
        n = 1
        for d in self.frame_shape:
            n *= d
        self.slot_bytes = (n + 255) // 256 * 256
        total = self.world * self.slot_bytes + self.FLAG_BYTES
        dev = torch.device("cuda", torch.cuda.current_device())
        # Every rank executes the same collectives whatever fails locally (a rank that raised before a collective would
        # leave the others hanging in it); errors are exchanged and raised by all ranks together.
        err, box = None, [None]
        if self.rank == dst:
            try:
                self.buf = torch.zeros(total, dtype=torch.uint8, device=dev)
                torch.cuda.synchronize()
                if self.world > 1:
                    from torch.multiprocessing.reductions import reduce_tensor
                    box = [reduce_tensor(self.buf)]                          # CUDA IPC handle of the allocation
            except Exception as exc:  # noqa: BLE001
                err = f"rank {self.rank}: {type(exc).__name__}: {exc}"
        if self.world > 1:
            dist.broadcast_object_list(box, src=dst)
            if self.rank != dst:
                try:
                    if box[0] is None:
                        raise RuntimeError("the destination rank could not export its buffer")
                    rebuild, args = box[0]
                    # Open the handle with OUR device current (argument 6 of rebuild_cuda_tensor is the device the storage
                    # is opened on): cudaIpcOpenMemHandle(..., cudaIpcMemLazyEnablePeerAccess) then maps dst's memory for
                    # this GPU.  Opened on dst's device index (what the exporter recorded) the mapping belongs to that
                    # device's context in this process and a kernel on our GPU faults on it (seen on 2 x B200).
                    args = list(args)
                    if not (len(args) >= 8 and isinstance(args[6], int)):
                        raise RuntimeError("unexpected CUDA IPC descriptor layout from torch.multiprocessing.reductions")
                    self.ops.peer_enable(args[6])
                    args[6] = dev.index
                    self.buf = rebuild(*args)                                # dst's memory, addressable from this GPU
                except Exception as exc:  # noqa: BLE001
                    err = f"rank {self.rank}: {type(exc).__name__}: {exc}"
            errs = [None] * self.world
            dist.all_gather_object(errs, err)
            errs = [e for e in errs if e]
            if errs:
                raise RuntimeError("PeerFrameGather unavailable: " + "; ".join(errs)[:300])
        elif err:
            raise RuntimeError(err)
        self.flags = self.buf[self.world * self.slot_bytes:].view(torch.int32)
        self.ready = self.flags[:self.world]
        self.consumed = self.flags[64:65]
        self.timed_out = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch = 0
        if self.world > 1:
            dist.barrier()

    def slot(self, r=None):
        r = self.rank if r is None else r
        n = 1
        for d in self.frame_shape:
            n *= d
        return self.buf[r * self.slot_bytes: r * self.slot_bytes + n].view(self.frame_shape)

    # -- writer side (every rank) ---------------------------------------------------------------
    def begin(self):
        """Start of a clip's frame output: returns this rank's slot; its first store is ordered after dst's release of
        the previous clip."""
        self.epoch += 1
        if self.epoch > 1:
            self.ops.peer_wait(self.consumed, self.epoch - 1, self.timeout_s, self.timed_out)
        return self.slot()

    def publish(self):
        self.ops.peer_signal(self.ready[self.rank:self.rank + 1], self.epoch)

    # -- reader side (dst) ----------------------------------------------------------------------
    def collect(self):
        assert self.rank == self.dst
        self.ops.peer_wait(self.ready, self.epoch, self.timeout_s, self.timed_out)
        n = 1
        for d in self.frame_shape:
            n *= d
        if self.slot_bytes == n:
            return self.buf[: self.world * n].view(self.world, *self.frame_shape)
        return torch.stack([self.slot(r) for r in range(self.world)])

    def release(self):
        assert self.rank == self.dst
        self.ops.peer_signal(self.consumed, self.epoch)

    def close(self):
        """Collective: importers drop their mapping of dst's buffer before dst frees it (otherwise the exporting process
        warns that it terminated with shared CUDA tensors outstanding)."""
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        if self.rank != self.dst:
            self.flags = self.ready = self.consumed = None
            self.buf = None
        if self.world > 1:
            dist.barrier()

    def check(self):
        """After a synchronize: raise if a wait gave up (a peer died or never published)."""
        v = int(self.timed_out.item())
        if v:
            raise RuntimeError(f"PeerFrameGather: timed out waiting for flag {v - 1} (epoch {self.epoch})")
