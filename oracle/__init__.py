"""ORACLE -- test infrastructure only.

A CPU/fp32 PyTorch restatement of the reference hot path (SVD denoise loop + MOFA-Adapter +
softsplat + Euler scheduler).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package (mofa_video_b200/) never does.

Parity status: the reference cannot be imported here (diffusers 0.24.0 and cupy are absent, SURVEY.md
§8c) and ships no golden vectors, so the diffusers-derived blocks are "parity unpinned"
(oracle/d24_blocks.py header).  The in-tree pieces that CAN be executed from /root/reference with
import stubs -- the Euler scheduler, the adapter's conditioning/first-frame encoders, the CMP modules, the
occlusion hourglass, and the UNet / ControlNetSDVModel / FlowControlNet definitions and forward graphs (with the
block classes bound to oracle/d24_blocks.py) -- are pinned by fixtures generated from the reference itself
(oracle/make_goldens.py -> tests/golden/).
"""
