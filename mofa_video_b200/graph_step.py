"""One denoise step of the Traj loop (/root/reference/MOFA-Video-Traj/pipeline/pipeline.py:447-511) as a replayable unit.

A step = MOFA-Adapter trunk + UNet + fused CFG / Euler / next-input kernel: ~950 kernel launches that the reference
(and round 1 of this engine) enqueue from Python every step.  Here the step is captured ONCE per pipeline and shape
into a CUDA graph and replayed 25 times per clip, clip after clip:

  * nothing in the captured launches depends on the step index: the timestep reaches `mofa_timestep_embedding` and
    (sigma, sigma_next) reach `mofa_cfg_euler_step_dev` through a small device buffer `cur` = [t] * B + [sigma, sigma_next]
    that is refilled (one 16-byte device copy, outside the graph) before each replay from a per-clip table;
  * every tensor the graph reads from outside itself is persistent storage owned by this object or by the networks
    (`Net.pbuf`): latents state, image latents, model input, the adapter's hoisted warped features, the collapsed
    cross-attention vectors and the added-time embedding -- a new clip overwrites their CONTENTS, never their addresses
    (the TMA descriptors inside the graph are built from raw addresses);
  * activations are allocated from the graph's private pool during capture and reused by every replay.

The host cost of a step drops from ~0.2 s of Python / ctypes enqueueing to one `cudaGraphLaunch`; the device work is
unchanged.  While per-kernel profiling is on (`lib.profile_start`, bench.py's roofline pass) or on the CPU test backend
the same body runs eagerly, so both paths execute identical code.
"""
import torch


class StepRunner:
    def __init__(self, ops, unet_net, ad_net, T, h, w, g_min, g_max, cond_scale, device):
        self.ops, self.unet_net, self.ad_net = ops, unet_net, ad_net
        self.T, self.h, self.w, self.hw = T, h, w, h * w
        self.g_min, self.g_max, self.cond_scale = float(g_min), float(g_max), float(cond_scale)
        self.device = torch.device(device)
        self.B = 2
        hw = self.hw
        self.lat_h = torch.empty(T, 4, hw, dtype=torch.float16, device=self.device)
        self.img_lat = torch.empty(2, 4, hw, dtype=torch.float16, device=self.device)
        self.next_in = torch.empty(2 * T * hw, 8, dtype=torch.float16, device=self.device)
        self.cur = torch.zeros(self.B + 2, dtype=torch.float32, device=self.device)
        self.table = None
        self.graph = None
        self.kernels_per_step = 0
        self.warm = False
        self.use_graph = self.device.type == "cuda"
        for n in (unet_net, ad_net):
            n.persistent = True     # their per-clip conditioning tensors keep their addresses from now on

    # ------------------------------------------------------------------ per clip
    def begin_clip(self, latents_h, image_latents, timesteps, sigmas):
        """latents_h [T,4,hw], image_latents [2,4,hw] (any float dtype); timesteps / sigmas: host lists (len n, n+1)."""
        self.lat_h.copy_(latents_h)
        self.img_lat.copy_(image_latents)
        n = len(timesteps)
        rows = [[float(timesteps[i])] * self.B + [float(sigmas[i]), float(sigmas[i + 1])] for i in range(n)]
        self.table = torch.tensor(rows, dtype=torch.float32).to(self.device, non_blocking=True)
        self.rebuild_input(float(sigmas[0]))

    def rebuild_input(self, sigma_next):
        """Model input of the next step from the current latents (loop prologue, or after a callback edited them)."""
        self.ops.cfg_euler_step(None, self.lat_h, self.img_lat, self.next_in, self.T, self.hw, self.g_min, self.g_max,
                                0.0, sigma_next)

    # ------------------------------------------------------------------ per step
    def _body(self):
        res, mid = self.ad_net.adapter_forward(self.next_in, self.cur[:self.B], self.h, self.w, self.cond_scale)
        noise = self.unet_net.unet_forward(self.next_in, self.cur[:self.B], self.h, self.w, res, mid)
        self.ops.cfg_euler_step_dev(noise, self.lat_h, self.img_lat, self.next_in, self.T, self.hw, self.g_min,
                                    self.g_max, self.cur[self.B:])

    def step(self, i):
        self.cur.copy_(self.table[i])
        ops = self.ops
        if not self.use_graph or ops.profiling():
            self._body()
            self.warm = True
            return
        if self.graph is None:
            if not self.warm:
                # first step ever: eager (one-time cudaFuncSetAttribute / driver entry-point lookups happen here)
                self._body()
                self.warm = True
                return
            n0 = ops.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._body()
            self.kernels_per_step = ops.launch_count() - n0
            ops.note_graph_replay(-self.kernels_per_step)   # the capture enqueued nothing: only replays execute kernels
            self.graph = g
        self.graph.replay()
        ops.note_graph_replay(self.kernels_per_step)


class HybridStepRunner(StepRunner):
    """The Hybrid step (/root/reference/MOFA-Video-Hybrid/pipeline/pipeline.py:432-507): landmark adapter + trajectory
    adapter on the same fused CFG input, their 12 + 1 residuals blended with the per-level nearest-resized mask, UNet,
    fused CFG / Euler -- captured and replayed like the Traj step.  `ad_net` = the face (landmark) adapter."""

    def __init__(self, ops, unet_net, face_net, drag_net, T, h, w, g_min, g_max, scale_ldmk, scale_traj, device):
        super().__init__(ops, unet_net, face_net, T, h, w, g_min, g_max, scale_ldmk, device)
        self.drag_net, self.scale_traj = drag_net, float(scale_traj)
        drag_net.persistent = True
        self.masks = {}          # rows of a residual -> fp16 mask [h_l * w_l] (persistent storage)

    def set_masks(self, by_rows):
        for rows, m in by_rows.items():
            if rows not in self.masks:
                self.masks[rows] = torch.empty_like(m)
            self.masks[rows].copy_(m)

    def _body(self):
        ops = self.ops
        t_dev = self.cur[:self.B]
        rf, mf = self.ad_net.adapter_forward(self.next_in, t_dev, self.h, self.w, self.cond_scale)
        rd, md = self.drag_net.adapter_forward(self.next_in, t_dev, self.h, self.w, self.scale_traj)
        for a, b in zip(rf + [mf], rd + [md]):
            m = self.masks[a.shape[0]]
            ops.mask_blend(a, b, m, a, period_rows=m.shape[0])           # face inside the mask, drag outside (:479-488)
        noise = self.unet_net.unet_forward(self.next_in, t_dev, self.h, self.w, rf, mf)
        ops.cfg_euler_step_dev(noise, self.lat_h, self.img_lat, self.next_in, self.T, self.hw, self.g_min, self.g_max,
                               self.cur[self.B:])
