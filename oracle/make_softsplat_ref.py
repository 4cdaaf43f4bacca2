"""ORACLE tooling (test infrastructure): pin the softsplat kernel -- the ONE native kernel the reference owns -- by
compiling and running the reference's own CUDA-C.

The reference builds its kernel at run time: `softsplat_func.forward` hands a CUDA-C template to `cuda_kernel()`
(/root/reference/MOFA-Video-Traj/models/softsplat.py:27-216), which bakes tensor sizes / strides into the text with
regular expressions, and CuPy compiles the result (:219-226).  CuPy is absent and this container has no GPU, so the
work is split (nothing from the reference is copied into the tracked tree):

  step 1 (HERE, needs /root/reference)        python -m oracle.make_softsplat_ref --build
      imports the reference's softsplat.py with a stub `cupy`, pulls the template out of `softsplat_func.forward`,
      runs the reference's OWN `cuda_kernel()` templating for every case shape below, adds a 6-line host launcher per
      case (grid = ceil(n / 512), block = 512, exactly :337-339) and compiles with nvcc for sm_100a into
      oracle/_ref/libsoftsplat_ref.so.  oracle/_ref/ is git-ignored but NOT gpurun-ignored: it travels to the GPU box.
  step 2 (GPU box)                            python -m oracle.make_softsplat_ref --run
      launches the reference kernel on the seeded inputs of every case (input = cat(x, ones) as softsplat.py:246 does,
      output pre-zeroed as :281) and writes the raw kernel outputs to gpurun_out/softsplat_ref_raw.pt.
  step 3 (HERE)                               python -m oracle.make_softsplat_ref --finish
      runs the reference's own Python wrapper `softsplat(x, flow, None, 'avg')` (:232-274) with `softsplat_func.apply`
      answered from the recorded kernel outputs (inputs checked bit-exactly), and commits the result as
      tests/golden/softsplat_ref.pt = {cases: [{x, flow, out_sum, out_avg}]}.

tests/test_oracle.py pins oracle/softsplat.py against that fixture on CPU; tests/test_softsplat_ref_gpu.py pins
`mofa_softsplat_avg` (through models.softsplat.softsplat) against it on the B200 and, when oracle/_ref is present,
also re-runs the reference kernel live.
"""
import argparse
import ctypes
import inspect
import os
import re
import subprocess
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FILE = "/root/reference/MOFA-Video-Traj/models/softsplat.py"
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
LIB = os.path.join(REF_DIR, "libsoftsplat_ref.so")
RAW = os.path.join(ROOT, "gpurun_out", "softsplat_ref_raw.pt")
GOLDEN = os.path.join(ROOT, "tests", "golden", "softsplat_ref.pt")

# (N, C, H, W, flow kind): C excludes the ones channel softsplat() appends for 'avg'
CASES = [
    (2, 8, 12, 20, "noise3"),       # fractional weights, collisions
    (1, 16, 36, 64, "rot"),         # the synthetic field of SURVEY 8d at the 1/16 pyramid level
    (2, 4, 9, 16, "big"),           # most targets out of bounds
    (1, 8, 12, 20, "nonfinite"),    # inf / nan flow entries are skipped (:301-302)
    (1, 8, 12, 20, "zero"),         # identity: x / (1 + 1e-7)
    (1, 4, 8, 8, "integer"),        # exact shifts, exact collisions at the border
]


def make_case(i):
    """Seeded inputs of case i: x fp32 [N,C,H,W] (fp16-representable), flow fp32 [N,2,H,W] (fp16-representable)."""
    N, C, H, W, kind = CASES[i]
    g = torch.Generator().manual_seed(100 + i)
    x = torch.randn(N, C, H, W, generator=g).half().float()
    if kind == "noise3":
        fl = torch.randn(N, 2, H, W, generator=g) * 3
    elif kind == "rot":
        ys = torch.arange(H, dtype=torch.float32)[:, None] / H
        xs = torch.arange(W, dtype=torch.float32)[None, :] / W
        A = 0.06 * min(H, W) * 4
        fl = torch.stack([torch.sin(6.2832 * ys) * torch.cos(6.2832 * xs) * A + 0 * xs,
                          torch.cos(6.2832 * ys) * torch.sin(6.2832 * xs) * A + 0 * ys], 0)[None]
        fl = fl + torch.randn(N, 2, H, W, generator=g)
    elif kind == "big":
        fl = torch.randn(N, 2, H, W, generator=g) * 12
    elif kind == "nonfinite":
        fl = torch.randn(N, 2, H, W, generator=g) * 2
        fl[0, 0, 1, 2] = float("inf")
        fl[0, 1, 3, 4] = float("nan")
        fl[0, 0, 5, 6] = float("-inf")
    elif kind == "zero":
        fl = torch.zeros(N, 2, H, W)
    elif kind == "integer":
        fl = torch.randint(-3, 4, (N, 2, H, W), generator=g).float()
    else:
        raise ValueError(kind)
    return x, fl.half().float()


# ------------------------------------------------------------------------------------------------ step 1
def _import_reference():
    """The reference module with `cupy` stubbed (only names touched at import / templating time)."""
    cp = types.ModuleType("cupy")
    cp.int32, cp.float32 = int, float
    cp.memoize = lambda **k: (lambda f: f)
    cp.cuda = types.SimpleNamespace(get_cuda_path=lambda: "/usr/local/cuda", compile_with_cache=None)
    sys.modules["cupy"] = cp
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_softsplat", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build():
    if not os.path.exists(REF_FILE):
        print(f"{REF_FILE} not present: oracle/_ref is only built in the builder container")
        return None
    ref = _import_reference()
    torch.cuda.get_device_name = lambda *a: "B200"       # cuda_kernel() only uses it as a cache key (:28-30)
    src = inspect.getsource(ref.softsplat_func.forward)
    m = re.search(r"cuda_kernel\('softsplat_out', '''(.*?)''', \{", src, re.S)
    assert m, "softsplat_out template not found in the reference"
    template = m.group(1)
    parts = ["#include <cuda_runtime.h>\n#include <assert.h>\n#include <math.h>\n"]
    for i, (N, C, H, W, _) in enumerate(CASES):
        ten_in = torch.empty(N, C + 1, H, W)
        ten_flow = torch.empty(N, 2, H, W)
        ten_out = torch.empty(N, C + 1, H, W)
        key = ref.cuda_kernel("softsplat_out", template, {"tenIn": ten_in, "tenFlow": ten_flow, "tenOut": ten_out})
        code = ref.objCudacache[key]["strKernel"]
        assert "{{" not in code and "SIZE_" not in code and "VALUE_" not in code and "OFFSET_" not in code
        code = code.replace("softsplat_out(", f"softsplat_out_{i}(")
        parts.append(code)
        parts.append(f"""
extern "C" int launch_softsplat_out_{i}(const float* tenIn, const float* tenFlow, float* tenOut, void* stream) {{
    const int n = {N * (C + 1) * H * W};
    softsplat_out_{i}<<<(n + 512 - 1) / 512, 512, 0, (cudaStream_t)stream>>>(n, tenIn, tenFlow, tenOut);
    return (int)cudaGetLastError();
}}
""")
    os.makedirs(REF_DIR, exist_ok=True)
    cu = os.path.join(REF_DIR, "softsplat_ref.cu")
    with open(cu, "w") as f:
        f.write("\n".join(parts))
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler",
           "-fPIC", "-o", LIB, cu]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed on the templated reference kernel")
    print(f"built {LIB} ({len(CASES)} templated instances of the reference's softsplat_out)")
    return LIB


# ------------------------------------------------------------------------------------------------ step 2
def run_reference_kernel(i, x, fl, lib=None):
    """Launch instance i of the reference kernel on cuda:0.  Returns the raw summation splat of cat(x, ones)."""
    lib = lib or ctypes.CDLL(LIB)
    fn = getattr(lib, f"launch_softsplat_out_{i}")
    fn.argtypes = [ctypes.c_void_p] * 4
    ten_in = torch.cat([x, x.new_ones(x.shape[0], 1, x.shape[2], x.shape[3])], 1).cuda().contiguous()   # :246
    ten_flow = fl.cuda().contiguous()
    ten_out = torch.zeros_like(ten_in)                                                                   # :281
    rc = fn(ten_in.data_ptr(), ten_flow.data_ptr(), ten_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0, f"reference kernel launch failed: {rc}"
    return ten_out.cpu()


def run():
    lib = ctypes.CDLL(LIB)
    raw = []
    for i in range(len(CASES)):
        x, fl = make_case(i)
        raw.append(run_reference_kernel(i, x, fl, lib))
    os.makedirs(os.path.dirname(RAW), exist_ok=True)
    torch.save({"raw": raw, "device": torch.cuda.get_device_name()}, RAW)
    print(f"wrote {RAW}: {[tuple(r.shape) for r in raw]}")


# ------------------------------------------------------------------------------------------------ step 3
def finish():
    ref = _import_reference()
    rec = torch.load(RAW)
    cases = []
    for i in range(len(CASES)):
        x, fl = make_case(i)
        raw = rec["raw"][i]

        class _Recorded:
            @staticmethod
            def apply(ten_in, ten_flow):
                want = torch.cat([x, x.new_ones(x.shape[0], 1, x.shape[2], x.shape[3])], 1)
                assert torch.equal(ten_in, want) and torch.equal(ten_flow.nan_to_num(7.0, 8.0, 9.0),
                                                                 fl.nan_to_num(7.0, 8.0, 9.0))
                return raw.clone()
        ref.softsplat_func = _Recorded           # the reference wrapper's only use of the kernel (:250)
        out_avg = ref.softsplat(x, fl, None, "avg")
        cases.append({"x": x.half(), "flow": fl.half(), "out_sum": raw, "out_avg": out_avg})
    torch.save({"cases": cases, "kinds": [c[4] for c in CASES], "device": rec.get("device"),
                "source": "reference CUDA-C softsplat_out (softsplat.py:285-335) templated by the reference's "
                          "cuda_kernel(), compiled with nvcc, run on the GPU box; avg by the reference's softsplat()"},
               GOLDEN)
    print(f"wrote {GOLDEN}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--finish", action="store_true")
    a = ap.parse_args()
    if a.build:
        build()
    if a.run:
        run()
    if a.finish:
        finish()
