"""The product's parameter enumeration (mofa_video_b200/synthetic.py) must equal the oracle modules' state
dicts key-for-key and shape-for-shape, and reproduce the published SVD-XT UNet size."""
import torch

from mofa_video_b200 import synthetic
from oracle import fixtures
from oracle.models import FlowControlNet, UNetSpatioTemporalConditionControlNetModel


def _shapes(sd):
    return {k: tuple(v.shape) for k, v in sd.items()}


def test_tiny_keys_match_oracle():
    cfg = dict(fixtures.TINY_CONFIG)
    _, su = synthetic.unet_state_dict(cfg)
    _, sa = synthetic.adapter_state_dict(cfg)
    ou = UNetSpatioTemporalConditionControlNetModel(**cfg).state_dict()
    oa = FlowControlNet(**cfg).state_dict()
    assert _shapes(su) == _shapes(ou)
    assert _shapes(sa) == _shapes(oa)


def test_full_size_param_count():
    with torch.device("meta"):
        ou = UNetSpatioTemporalConditionControlNetModel()
        oa = FlowControlNet()
    n_u = sum(p.numel() for p in ou.parameters())
    assert n_u == 1_524_623_082  # published SVD-XT UNet size
    assert abs(sum(p.numel() for p in oa.parameters()) - 694.3e6) < 0.1e6


def test_ldmk_adapter_keys_match_oracle():
    from oracle.keypoint import FlowControlNetLdmk
    cfg = dict(fixtures.TINY_CONFIG)
    _, sl = synthetic.ldmk_adapter_state_dict(cfg)
    assert _shapes(sl) == _shapes(FlowControlNetLdmk(**cfg).state_dict())
    with torch.device("meta"):
        full = FlowControlNetLdmk()
    _, sf = synthetic.ldmk_adapter_state_dict({"num_frames": 2}, dtype=torch.float16)
    assert {k: tuple(v.shape) for k, v in sf.items()} == {k: tuple(v.shape) for k, v in full.state_dict().items()}
