"""Control-signal rasterisation: sparse motion hints -> dense (flow, mask) planes.
CPU: oracle/sparse_flow.py and the product wrappers (PyTorch statements of the kernels) vs the golden produced by the
reference's own functions (bit-exact), error behaviour.  GPU: mofa_sparse_hints vs the same golden, bit-exact."""
import os

import numpy as np
import pytest
import torch

import ref_ops
from mofa_video_b200.utils import sparse_flow as sf
from oracle import sparse_flow as osf
from oracle.make_goldens import sparse_flow_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sparse_flow_ref.pt")


def _check(fn_traj, fn_ldmk, dev):
    g, c = torch.load(GOLD), sparse_flow_cases()
    for back in (False, True):
        f, m = fn_traj(c["points"], c["n_steps"], c["H"], c["W"], is_backward_flow=back)
        f, m = torch.as_tensor(f).double().cpu(), torch.as_tensor(m).double().cpu()
        assert torch.equal(f, g[f"traj_flow_back{int(back)}"]) and torch.equal(m, g[f"traj_mask_back{int(back)}"])
    assert g["traj_mask_back0"].max() == 2.0                       # the colliding tracks were added, not overwritten
    f, m = fn_ldmk(c["landmarks"].to(dev), c["H"], c["W"], c["t"])
    assert torch.equal(f.cpu(), g["ldmk_flow"]) and torch.equal(m.cpu(), g["ldmk_mask"])
    assert m.dtype == torch.uint8 and f.shape == (2, 3, 2, c["H"], c["W"])


def test_oracle_and_host_wrappers_match_reference_functions():
    _check(osf.get_sparseflow_and_mask_forward, osf.get_sparse_flow, "cpu")
    _check(lambda *a, **k: sf.get_sparseflow_and_mask_forward(*a, device="cpu", ops=ref_ops, **k),
           lambda *a: sf.get_sparse_flow(*a, ops=ref_ops), "cpu")
    c = sparse_flow_cases()
    bad = c["points"].copy()
    bad[0, 0, 1] = c["H"] + 3                                       # numpy: IndexError (run_gradio.py:79)
    with pytest.raises(IndexError):
        sf.get_sparseflow_and_mask_forward(bad, c["n_steps"], c["H"], c["W"], device="cpu", ops=ref_ops)
    with pytest.raises(IndexError):
        osf.get_sparseflow_and_mask_forward(bad, c["n_steps"], c["H"], c["W"])
    f, m = sf.get_sparseflow_and_mask_forward(np.zeros((0, 6, 2)), 5, 8, 8, device="cpu", ops=ref_ops)
    assert float(f.abs().sum()) == 0 and float(m.sum()) == 0


@pytest.mark.gpu
def test_sparse_hints_kernel_bit_exact():
    _check(sf.get_sparseflow_and_mask_forward, sf.get_sparse_flow, "cuda")
    # negative start indices wrap like numpy; float64 landmarks keep their dtype
    c = sparse_flow_cases()
    pts = c["points"].copy()
    pts[1, 0] = [-2.5, -1.2]
    a, b = osf.get_sparseflow_and_mask_forward(pts, c["n_steps"], c["H"], c["W"])
    f, m = sf.get_sparseflow_and_mask_forward(pts, c["n_steps"], c["H"], c["W"])
    assert torch.equal(f.double().cpu(), torch.from_numpy(a)) and torch.equal(m.double().cpu(), torch.from_numpy(b))
    lm = c["landmarks"].double()
    a, b = osf.get_sparse_flow(lm, c["H"], c["W"], c["t"])
    f, m = sf.get_sparse_flow(lm.cuda(), c["H"], c["W"], c["t"])
    assert f.dtype == torch.float64 and torch.equal(f.cpu(), a) and torch.equal(m.cpu(), b)
    # full size (384 x 384, 24 steps, 32 tracks) as T/run_gradio.py uses it
    rng = np.random.default_rng(1)
    pts = np.zeros((32, 25, 2))
    pts[:, 0] = rng.uniform(0, 383, (32, 2))
    pts[:, 1:] = pts[:, 0:1] + np.cumsum(rng.normal(0, 3, (32, 24, 2)), axis=1)
    a, b = osf.get_sparseflow_and_mask_forward(pts, 24, 384, 384)
    f, m = sf.get_sparseflow_and_mask_forward(pts, 24, 384, 384)
    assert torch.equal(f.double().cpu(), torch.from_numpy(a)) and torch.equal(m.double().cpu(), torch.from_numpy(b))
