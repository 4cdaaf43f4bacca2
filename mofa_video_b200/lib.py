"""ctypes binding of libmofa_b200.so (C ABI declared in include/mofa_b200.h).

Every wrapper takes torch CUDA tensors, passes raw data_ptr()s and the current torch stream -- the
same calling convention the reference uses for its one native kernel
(/root/reference/MOFA-Video-Traj/models/softsplat.py:340-345).  There is NO fallback: if the shared
library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmofa_b200.so")

A_LINEAR, A_CONV3X3, A_TEMPORAL3 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3, 4
ACT_GELU, ACT_QUICK_GELU = 6, 7

EXPORTS = [
    "mofa_last_error", "mofa_version", "mofa_launch_count", "mofa_launch_count_reset", "mofa_gemm",
    "mofa_attn_spatial", "mofa_attn_temporal", "mofa_groupnorm", "mofa_layernorm", "mofa_axpy_bcast",
    "mofa_im2col3x3", "mofa_upsample2x", "mofa_nchw_to_nhwc", "mofa_nhwc_to_nchw", "mofa_linear_small",
    "mofa_timestep_embedding", "mofa_softsplat_avg", "mofa_cfg_euler_step", "mofa_softmax_rows",
    "mofa_vae_time_conv_out", "mofa_im2col", "mofa_pool2d", "mofa_resize_bilinear_ac", "mofa_cmp_fuser",
    "mofa_copy_cols", "mofa_flow_pyramid", "mofa_mask_blend", "mofa_downsample_nearest", "mofa_flow_post",
    "mofa_resize_antialias", "mofa_cfg_euler_step_dev", "mofa_sparse_hints", "mofa_peer_enable", "mofa_peer_signal", "mofa_peer_wait",
    "mofa_attn_small", "mofa_attn_small_temporal", "mofa_ff_geglu", "mofa_ff_debug_dump",
]


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("mode", ctypes.c_int32), ("act", ctypes.c_int32),
        ("a", ctypes.c_void_p), ("a2", ctypes.c_void_p), ("w", ctypes.c_void_p), ("out", ctypes.c_void_p),
        ("ldc", ctypes.c_int64),
        ("M", ctypes.c_int64), ("K", ctypes.c_int64), ("K1", ctypes.c_int64), ("lda", ctypes.c_int64),
        ("lda2", ctypes.c_int64),
        ("n_img", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("C", ctypes.c_int32),
        ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("HW", ctypes.c_int32),
        ("N", ctypes.c_int32), ("bn", ctypes.c_int32),
        ("bias", ctypes.c_void_p), ("rowbias", ctypes.c_void_p), ("ld_rowbias", ctypes.c_int64),
        ("rows_per_group", ctypes.c_int64), ("rowbias_mod", ctypes.c_int64),
        ("res1", ctypes.c_void_p), ("ldr1", ctypes.c_int64), ("res2", ctypes.c_void_p), ("ldr2", ctypes.c_int64),
        ("alpha", ctypes.c_float), ("beta1", ctypes.c_float), ("beta2", ctypes.c_float),
        ("max_ctas", ctypes.c_int32), ("dilation", ctypes.c_int32), ("ksize", ctypes.c_int32),
        ("gn_stats", ctypes.c_void_p), ("gn_rows_per_stat", ctypes.c_int64), ("gn_groups", ctypes.c_int32),
        ("gn_cpg", ctypes.c_int32), ("gn_c_off", ctypes.c_int32), ("gn_pad", ctypes.c_int32),
    ]


_lib = None


def load():
    """Load the shared library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; "
            "g.build()'`). mofa_video_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mofa_last_error.restype = ctypes.c_char_p
    lib.mofa_launch_count.restype = ctypes.c_int64
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    lib.mofa_gemm.argtypes = [ctypes.POINTER(GemmArgs), vp]
    lib.mofa_attn_spatial.argtypes = [vp, vp, i32, i32, i32, f32, vp]
    lib.mofa_attn_temporal.argtypes = [vp, vp, i32, i32, i32, i32, f32, vp]
    lib.mofa_groupnorm.argtypes = [vp, i32, vp, i32, vp, vp, vp, i64, i64, i32, f32, i32, vp, vp]
    lib.mofa_layernorm.argtypes = [vp, vp, vp, vp, i64, i32, f32, vp, i64, i64, vp, vp]
    lib.mofa_axpy_bcast.argtypes = [vp, vp, vp, i64, i64, f32, vp]
    lib.mofa_im2col3x3.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.mofa_upsample2x.argtypes = [vp, vp, i32, i32, i32, i32, vp]
    lib.mofa_nchw_to_nhwc.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.mofa_nhwc_to_nchw.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.mofa_linear_small.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.mofa_timestep_embedding.argtypes = [vp, vp, i32, i32, vp]
    lib.mofa_softsplat_avg.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    lib.mofa_cfg_euler_step.argtypes = [vp, vp, vp, vp, i32, i32, f32, f32, f32, f32, vp]
    lib.mofa_cfg_euler_step_dev.argtypes = [vp, vp, vp, vp, i32, i32, f32, f32, vp, vp]
    lib.mofa_sparse_hints.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.mofa_peer_enable.argtypes = [i32]
    lib.mofa_peer_signal.argtypes = [vp, ctypes.c_uint32, vp]
    lib.mofa_peer_wait.argtypes = [vp, i32, ctypes.c_uint32, ctypes.c_double, vp, vp]
    lib.mofa_ff_geglu.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, vp, i64, vp, i64, f32, f32, f32, vp]
    lib.mofa_attn_small.argtypes = [vp, vp, i32, i32, i32, i32, f32, vp]
    lib.mofa_attn_small_temporal.argtypes = [vp, vp, i32, i32, i32, i32, i32, f32, vp]
    lib.mofa_softmax_rows.argtypes = [vp, i64, i32, i64, vp]
    lib.mofa_vae_time_conv_out.argtypes = [vp, vp, vp, vp, vp, i32, i64, vp]
    lib.mofa_im2col.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.mofa_pool2d.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.mofa_resize_bilinear_ac.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.mofa_cmp_fuser.argtypes = [vp, vp, i64, i32, f32, vp]
    lib.mofa_copy_cols.argtypes = [vp, vp, i64, i32, i64, i32, i32, vp]
    lib.mofa_flow_pyramid.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]
    lib.mofa_mask_blend.argtypes = [vp, vp, vp, vp, i64, i32, i64, vp]
    lib.mofa_downsample_nearest.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    lib.mofa_flow_post.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]
    lib.mofa_resize_antialias.argtypes = [vp, vp, i32, i32, i32, i32, i32, vp]
    _lib = lib
    return lib


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {load().mofa_last_error().decode()}")


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _chk_h(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float16 and t.is_contiguous(), (t.dtype, t.shape, t.is_contiguous())


_graph_launches = 0  # kernels executed by CUDA-graph replays (the C counter only sees launches enqueued by the host)


def note_graph_replay(n_kernels):
    global _graph_launches
    _graph_launches += int(n_kernels)


def launch_count():
    return int(load().mofa_launch_count()) + _graph_launches


# optional per-launch device timing (CUDA events on the launching stream) used by bench.py's roofline line
_prof = None


def profile_start():
    global _prof
    _prof = []


def profile_stop(detail=False):
    """-> {kind: {"ms": device time, "work": algorithmic FLOPs (or bytes), "launches": n}}; detail=True keys the
    same aggregates by (kind + shape/epilogue tag) instead."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for kind, work, e0, e1, tag in rec or []:
        d = out.setdefault(f"{kind} {tag}" if detail else kind, {"ms": 0.0, "work": 0.0, "launches": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["work"] += work
        d["launches"] += 1
    return out


class _Timed:
    def __init__(self, kind, work, tag=""):
        self.kind, self.work, self.tag = kind, work, tag

    def __enter__(self):
        if _prof is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _prof is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _prof.append((self.kind, self.work, self.e0, e1, self.tag))
        return False


def launch_count_reset():
    global _graph_launches
    _graph_launches = 0
    load().mofa_launch_count_reset()


def profiling():
    return _prof is not None


def pick_bn(n, geglu=False):
    """N tile for mofa_gemm: the widest tile (<=256) that wastes the least of N."""
    if geglu:
        for bn in (256, 128):
            if n % bn == 0:
                return bn
        raise ValueError(f"GEGLU N={n} has no valid tile")
    if n <= 256:
        return max(16, (n + 15) // 16 * 16)
    # multiples of 64 so the 64-column TMA-store chunks never straddle tiles; the last tile of a row may be
    # narrower (its MMA N shrinks), so prefer the split whose remainder is not a sliver
    for bn in (256, 192, 128):
        rem = n % bn
        if rem == 0 or rem >= 128:
            return bn
    return 256


def gemm(mode, a, w, out, *, N, bn=None, act=ACT_NONE, a2=None, M=0, K=0, K1=0, lda=0, lda2=0, n_img=0, H=0, W=0,
         C=0, B=0, T=0, HW=0, ldc=None, bias=None, rowbias=None, rows_per_group=1, rowbias_mod=0, res1=None, res2=None,
         alpha=1.0, beta1=1.0, beta2=1.0, max_ctas=0, dilation=1, ksize=3, c_off=0, gn_stats=None,
         gn_rows_per_stat=0, gn_groups=32, gn_cpg=0, gn_c_off=0):
    """c_off: first output column inside rows of width ldc (in-place channel concat; multiple of 8).
    gn_stats (fp32 [rows / gn_rows_per_stat, gn_groups, 2]): GroupNorm statistics of this output, accumulated by the
    epilogue for the GroupNorm that consumes it (gn_cpg channels per group; default N_out / gn_groups)."""
    lib = load()
    _chk_h(a, a2, w, out, bias, res1, res2)
    if rowbias is not None:  # may be a column slice of a wider [groups, total] matrix
        assert rowbias.is_cuda and rowbias.dtype == torch.float16 and rowbias.stride(-1) == 1
    g = GemmArgs()
    g.mode, g.act = mode, act
    g.a, g.a2, g.w = a.data_ptr(), (a2.data_ptr() if a2 is not None else None), w.data_ptr()
    g.out = out.data_ptr() + 2 * c_off
    n_out = N // 2 if act == ACT_GEGLU else N
    g.ldc = ldc if ldc is not None else n_out
    g.M, g.K, g.K1, g.lda, g.lda2 = M, K, K1, lda, lda2
    g.n_img, g.H, g.W, g.C, g.B, g.T, g.HW = n_img, H, W, C, B, T, HW
    g.N = N
    g.bn = bn if bn is not None else pick_bn(N, act == ACT_GEGLU)
    g.bias = bias.data_ptr() if bias is not None else None
    if rowbias is not None:
        g.rowbias, g.ld_rowbias, g.rows_per_group = rowbias.data_ptr(), rowbias.stride(0), rows_per_group
        g.rowbias_mod = rowbias_mod
    else:
        g.rowbias, g.ld_rowbias, g.rows_per_group = None, 0, 1
    g.res1 = res1.data_ptr() if res1 is not None else None
    g.ldr1 = res1.shape[-1] if res1 is not None else 0
    g.res2 = res2.data_ptr() if res2 is not None else None
    g.ldr2 = res2.shape[-1] if res2 is not None else 0
    g.alpha, g.beta1, g.beta2 = alpha, beta1, beta2
    g.max_ctas = max_ctas
    g.dilation = dilation
    g.ksize = ksize
    if gn_stats is not None:
        assert gn_stats.dtype == torch.float32 and gn_stats.is_cuda and gn_rows_per_stat > 0
        g.gn_stats, g.gn_rows_per_stat, g.gn_groups = gn_stats.data_ptr(), gn_rows_per_stat, gn_groups
        g.gn_cpg = gn_cpg if gn_cpg else n_out // gn_groups
        g.gn_c_off = gn_c_off
    else:
        g.gn_stats = None
    if _prof is not None:
        if mode == A_LINEAR:
            kind, work = "gemm_linear", 2.0 * M * N * K
        elif mode == A_CONV3X3:
            kind, work = "gemm_conv3x3", 2.0 * n_img * H * W * N * ksize * ksize * C
        else:
            kind, work = "gemm_temporal3", 2.0 * B * T * HW * N * 3 * C
        rows = M or n_img * H * W or B * T * HW
        tag = (f"M={rows} N={N} K={K or C} act={act} a2={int(a2 is not None)} rb={int(rowbias is not None)} "
               f"res={int(res1 is not None) + int(res2 is not None)}")
        with _Timed(kind, work, tag):
            _check(lib.mofa_gemm(ctypes.byref(g), _stream()), "mofa_gemm")
        return out
    _check(lib.mofa_gemm(ctypes.byref(g), _stream()), "mofa_gemm")
    return out


def linear(a, w, out, **kw):
    """a [M, K] (row-major, last dim contiguous), w [N, K]."""
    M, K = a.shape[0], a.shape[1]
    return gemm(A_LINEAR, a, w, out, N=w.shape[0], M=M, K=K if "K" not in kw else kw.pop("K"), lda=a.stride(0), **kw)


def attn_spatial(qkv, out, frames, L, heads, scale):
    _chk_h(qkv, out)
    with _Timed("attn_spatial", 4.0 * frames * heads * L * L * 64, f"L={L} heads={heads}"):
        _check(load().mofa_attn_spatial(_p(qkv), _p(out), frames, L, heads, scale, _stream()), "mofa_attn_spatial")
    return out


# Widest FeedForward the engine routes through the fused kernel.  The kernel handles C <= 320 (tensor memory: the output
# accumulator's C columns + 128 + 64 must fit 512) and is correct, but with its N = 128 first MMA and single S buffer it
# is no faster in the step than the two pipelined GEMMs, so it is opt-in: MOFA_FF_FUSED=1.
FF_FUSED_LIMIT_C = 320
FF_FUSED_MAX_C = FF_FUSED_LIMIT_C if os.environ.get("MOFA_FF_FUSED", "0") == "1" else 0


def ff_geglu(x, w1_packed, b1_packed, w2, b2, out, res1=None, res2=None, alpha=1.0, beta1=1.0, beta2=1.0):
    """Fused GEGLU FeedForward (mofa_ff_geglu): x [M, C] -> out [M, C]; w1_packed / b1_packed in the bn = 128 GEGLU packing."""
    _chk_h(x, w1_packed, b1_packed, w2, b2, out, res1, res2)
    M, C = x.shape
    hidden = w2.shape[1]
    assert w1_packed.shape == (2 * hidden, C) and w2.shape[0] == C and out.shape == (M, C)
    work = 2.0 * M * C * hidden * 3
    with _Timed("ff_geglu_fused", work, f"M={M} C={C} hidden={hidden} res={int(res1 is not None) + int(res2 is not None)}"):
        _check(load().mofa_ff_geglu(_p(x), _p(w1_packed), _p(b1_packed), _p(w2), _p(b2), _p(out), M, C, hidden, _p(res1),
                                    res1.shape[-1] if res1 is not None else 0, _p(res2),
                                    res2.shape[-1] if res2 is not None else 0, alpha, beta1, beta2, _stream()),
               "mofa_ff_geglu")
    return out


def attn_small(qkv, out, n_seq, L, heads, head_dim, scale):
    """qkv [n_seq * L, 3 * heads * head_dim] -> out [n_seq * L, heads * head_dim] (CLIP image encoder)."""
    _chk_h(qkv, out)
    _check(load().mofa_attn_small(_p(qkv), _p(out), n_seq, L, heads, head_dim, scale, _stream()), "mofa_attn_small")
    return out


def attn_small_temporal(qkv, out, B, T, HW, heads, head_dim, scale):
    _chk_h(qkv, out)
    _check(load().mofa_attn_small_temporal(_p(qkv), _p(out), B, T, HW, heads, head_dim, scale, _stream()),
           "mofa_attn_small_temporal")
    return out


def attn_temporal(qkv, out, B, T, HW, heads, scale):
    _chk_h(qkv, out)
    _check(load().mofa_attn_temporal(_p(qkv), _p(out), B, T, HW, heads, scale, _stream()), "mofa_attn_temporal")
    return out


def groupnorm(x1, gamma, beta, out, rows_per_stat, eps, silu, stats, x2=None, groups=32, stats_ready=False):
    """stats_ready: `stats` already holds (sum, sum of squares) per (statistic, group) -- accumulated by the producing
    GEMM's epilogue (gemm(..., gn_stats=)) -- so only the apply pass runs."""
    _chk_h(x1, x2, gamma, beta, out)
    rows = x1.shape[0]
    C1 = x1.shape[1]
    C2 = x2.shape[1] if x2 is not None else 0
    assert stats.dtype == torch.float32 and stats.numel() >= (rows // rows_per_stat) * groups * 2
    _check(load().mofa_groupnorm(_p(x1), C1, _p(x2), C2, _p(gamma), _p(beta), _p(out), rows, rows_per_stat, groups,
                                 eps, (1 if silu else 0) | (2 if stats_ready else 0), _p(stats), _stream()),
           "mofa_groupnorm")
    return out


def layernorm(x, gamma, beta, out, eps=1e-5, add=None, rows_per_group=1, add_period=1, sum_out=None):
    _chk_h(x, gamma, beta, out, add, sum_out)
    _check(load().mofa_layernorm(_p(x), _p(gamma), _p(beta), _p(out), x.shape[0], x.shape[1], eps, _p(add),
                                 rows_per_group, add_period, _p(sum_out), _stream()), "mofa_layernorm")
    return out


def axpy_bcast(x, y, out, scale=1.0):
    _chk_h(x, y, out)
    _check(load().mofa_axpy_bcast(_p(x), _p(y), _p(out), x.numel(), y.numel(), scale, _stream()), "mofa_axpy_bcast")
    return out


def im2col3x3(x, out, n_img, H, W, C, stride, Kpad):
    _chk_h(x, out)
    _check(load().mofa_im2col3x3(_p(x), _p(out), n_img, H, W, C, stride, Kpad, _stream()), "mofa_im2col3x3")
    return out


def upsample2x(x, out, n_img, H, W, C):
    _chk_h(x, out)
    _check(load().mofa_upsample2x(_p(x), _p(out), n_img, H, W, C, _stream()), "mofa_upsample2x")
    return out


def nchw_to_nhwc(x, out, n_img, C, HW, ldo=None, c_off=0):
    _chk_h(x, out)
    _check(load().mofa_nchw_to_nhwc(_p(x), _p(out), n_img, C, HW, ldo if ldo is not None else C, c_off, _stream()),
           "mofa_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, out, n_img, C, HW, ldi=None, c_off=0):
    _chk_h(x, out)
    _check(load().mofa_nhwc_to_nchw(_p(x), _p(out), n_img, C, HW, ldi if ldi is not None else C, c_off, _stream()),
           "mofa_nhwc_to_nchw")
    return out


def linear_small(a, w, bias, out, act_in=0, act_out=0):
    _chk_h(a, w, bias, out)
    M, K = a.shape
    N = w.shape[0]
    _check(load().mofa_linear_small(_p(a), _p(w), _p(bias), _p(out), M, N, K, act_in, act_out, _stream()),
           "mofa_linear_small")
    return out


def timestep_embedding(t, out, dim):
    assert t.dtype == torch.float32 and t.is_cuda
    _chk_h(out)
    _check(load().mofa_timestep_embedding(_p(t), _p(out), t.numel(), dim, _stream()), "mofa_timestep_embedding")
    return out


def softsplat_avg(feat, flow, acc, wsum, out, F, hs, ws, C, Hf, Wf):
    _chk_h(feat, flow, out)
    assert acc.dtype == torch.float32 and wsum.dtype == torch.float32
    _check(load().mofa_softsplat_avg(_p(feat), _p(flow), _p(acc), _p(wsum), _p(out), F, hs, ws, C, Hf, Wf, _stream()),
           "mofa_softsplat_avg")
    return out


def cfg_euler_step(noise, latents_h, image_latents, next_in, T, HW, g_min, g_max, sigma, sigma_next):
    _chk_h(noise, latents_h, image_latents, next_in)
    _check(load().mofa_cfg_euler_step(_p(noise), _p(latents_h), _p(image_latents), _p(next_in), T, HW, g_min, g_max,
                                      sigma, sigma_next, _stream()), "mofa_cfg_euler_step")
    return next_in


def cfg_euler_step_dev(noise, latents_h, image_latents, next_in, T, HW, g_min, g_max, sigmas):
    """cfg_euler_step with (sigma, sigma_next) = sigmas[0:2], a float32 device tensor read at execution time."""
    _chk_h(noise, latents_h, image_latents, next_in)
    assert sigmas.is_cuda and sigmas.dtype == torch.float32 and sigmas.numel() >= 2
    _check(load().mofa_cfg_euler_step_dev(_p(noise), _p(latents_h), _p(image_latents), _p(next_in), T, HW, g_min, g_max,
                                          _p(sigmas), _stream()), "mofa_cfg_euler_step_dev")
    return next_in


def softmax_rows(x, L=None):
    """In-place softmax over the rows of a fp16 matrix [rows, ld] (first L columns)."""
    _chk_h(x)
    rows, ld = x.shape
    _check(load().mofa_softmax_rows(_p(x), rows, L if L is not None else ld, ld, _stream()), "mofa_softmax_rows")
    return x


def vae_time_conv_out(y, w, b, out_f32, out_u8, T, HW):
    _chk_h(y)
    assert w.dtype == torch.float32 and b.dtype == torch.float32
    _check(load().mofa_vae_time_conv_out(_p(y), _p(w), _p(b), _p(out_f32), _p(out_u8), T, HW, _stream()),
           "mofa_vae_time_conv_out")


def conv_out_size(n, k, stride, pad, dil=1):
    """pad < 0 means asymmetric padding: 0 before, -pad after (mofa_im2col)."""
    return (n + (-pad if pad < 0 else 2 * pad) - dil * (k - 1) - 1) // stride + 1


def im2col(x, out, n_img, H, W, C, ksize, stride, pad, dilation, Kpad):
    _chk_h(x, out)
    _check(load().mofa_im2col(_p(x), _p(out), n_img, H, W, C, ksize, stride, pad, dilation, Kpad, _stream()),
           "mofa_im2col")
    return out


def pool2d(x, out, n_img, H, W, C, ksize, stride, pad, mode):
    """mode 0 = max, 1 = average."""
    _chk_h(x, out)
    _check(load().mofa_pool2d(_p(x), _p(out), n_img, H, W, C, ksize, stride, pad, mode, _stream()), "mofa_pool2d")
    return out


def resize_bilinear_ac(x, out, n_img, H, W, C, Ho, Wo, ldo=None, c_off=0):
    _chk_h(x, out)
    _check(load().mofa_resize_bilinear_ac(_p(x), _p(out), n_img, H, W, C, Ho, Wo, ldo if ldo is not None else C,
                                          c_off, _stream()), "mofa_resize_bilinear_ac")
    return out


def cmp_fuser(logits, flow, nbins=99, fmax=50.0):
    _chk_h(logits, flow)
    _check(load().mofa_cmp_fuser(_p(logits), _p(flow), logits.shape[0], nbins, fmax, _stream()), "mofa_cmp_fuser")
    return flow


def copy_cols(src, dst, rows, C, period_rows, ldo, c_off):
    _chk_h(src, dst)
    _check(load().mofa_copy_cols(_p(src), _p(dst), rows, C, period_rows, ldo, c_off, _stream()), "mofa_copy_cols")
    return dst


def flow_pyramid(flow, out, F, hs, ws, Hf, Wf, ldo, c_off):
    _chk_h(flow, out)
    _check(load().mofa_flow_pyramid(_p(flow), _p(out), F, hs, ws, Hf, Wf, ldo, c_off, _stream()), "mofa_flow_pyramid")
    return out


def mask_blend(a, b, mask, out, period_rows=None):
    _chk_h(a, b, mask, out)
    rows, C = a.shape
    _check(load().mofa_mask_blend(_p(a), _p(b), _p(mask), _p(out), rows, C,
                                  period_rows if period_rows is not None else rows, _stream()), "mofa_mask_blend")
    return out


def downsample_nearest(x, out, n_img, H, W, C, s):
    _chk_h(x, out)
    _check(load().mofa_downsample_nearest(_p(x), _p(out), n_img, H, W, C, s, _stream()), "mofa_downsample_nearest")
    return out


def flow_post(flow_in, out, F, Hs, Ws, H, W, brush=None, flow_out=None):
    """Fused drag-flow post-processing (brush mask, nearest resize + rescale, in/out merge) on fp16 NCHW flows."""
    _chk_h(flow_in, out, brush, flow_out)
    _check(load().mofa_flow_post(_p(flow_in), _p(brush), _p(flow_out), _p(out), F, Hs, Ws, H, W, _stream()),
           "mofa_flow_post")
    return out


def sparse_hints_add(pts, flow, mask, sign=1):
    """mode 0: pts float64 [K, Tn, 2] on the device; flow fp32 [Tn-1, H, W, 2], mask fp32 [Tn-1, H, W] (zero-filled here)."""
    K, Tn, _ = pts.shape
    n, H, W, _ = flow.shape
    assert pts.is_cuda and pts.dtype == torch.float64 and pts.is_contiguous() and n == Tn - 1
    assert flow.dtype == mask.dtype == torch.float32 and flow.is_contiguous() and mask.is_contiguous()
    _check(load().mofa_sparse_hints(_p(pts), 0, 1, 1, Tn, K, H, W, sign, _p(flow), _p(mask), None, _stream()),
           "mofa_sparse_hints")
    return flow, mask


def sparse_hints_assign(landmarks, flow, mask, owner):
    """mode 1: landmarks [B, Tn, K, 2] fp32 / fp64; flow [B, Tn-1, 2, H, W] same dtype, mask uint8 same shape,
    owner int32 [B, Tn-1, H, W] workspace."""
    B, Tn, K, _ = landmarks.shape
    H, W = flow.shape[-2:]
    assert landmarks.is_cuda and landmarks.is_contiguous() and landmarks.dtype in (torch.float32, torch.float64)
    assert flow.dtype == landmarks.dtype and mask.dtype == torch.uint8 and owner.dtype == torch.int32
    _check(load().mofa_sparse_hints(_p(landmarks), 1, int(landmarks.dtype == torch.float64), B, Tn, K, H, W, 1,
                                    _p(flow), _p(mask), _p(owner), _stream()), "mofa_sparse_hints")
    return flow, mask


def peer_enable(peer_device):
    _check(load().mofa_peer_enable(int(peer_device)), "mofa_peer_enable")


def peer_signal(flag, value):
    """flag: one int32 element (local or IPC-mapped peer memory)."""
    assert flag.dtype == torch.int32 and flag.numel() >= 1
    _check(load().mofa_peer_signal(_p(flag), int(value) & 0xffffffff, _stream()), "mofa_peer_signal")


def peer_wait(flags, value, timeout_s=30.0, timed_out=None):
    assert flags.dtype == torch.int32 and flags.is_contiguous()
    _check(load().mofa_peer_wait(_p(flags), flags.numel(), int(value) & 0xffffffff, float(timeout_s), _p(timed_out),
                                 _stream()), "mofa_peer_wait")


def resize_antialias(img, out):
    """img fp32 [N, C, H, W] -> out fp32 [N, C, Ho, Wo]: Gaussian pre-blur + bicubic(align_corners=True) (CLIP input)."""
    assert img.is_cuda and out.is_cuda and img.dtype == out.dtype == torch.float32
    assert img.is_contiguous() and out.is_contiguous()
    n, c, H, W = img.shape
    _check(load().mofa_resize_antialias(_p(img), _p(out), n * c, H, W, out.shape[-2], out.shape[-1], _stream()),
           "mofa_resize_antialias")
    return out
