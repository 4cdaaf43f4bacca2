"""-m "not gpu": host logic of the native VAE temporal decoder (packing, attention-as-GEMMs, fused tail) with the
kernels replaced by their plain-PyTorch statements, against the fp32 PyTorch TemporalDecoder module."""
import torch

import ref_ops
from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
from mofa_video_b200.vae_engine import NativeTemporalDecoderVAE


def make_vae(seed=0):
    torch.manual_seed(seed)
    vae = AutoencoderKLTemporalDecoder(block_out_channels=(64, 64, 128, 128)).eval()
    with torch.no_grad():
        for n, p in vae.named_parameters():
            if n.endswith("mix_factor"):
                p.fill_(0.3)
            p.copy_(p.half().float())
    return vae


def test_native_decoder_matches_torch_module():
    vae = make_vae()
    g = torch.Generator().manual_seed(1)
    z = torch.randn(3, 4, 4, 6, generator=g).half().float()
    with torch.no_grad():
        ref = vae.decode(z, num_frames=3).sample
    nat = NativeTemporalDecoderVAE(vae, ops=ref_ops, device="cpu")
    out = nat.decode(z, num_frames=3).sample
    assert out.shape == ref.shape == (3, 3, 32, 48)
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, err
    u8 = nat.decode_uint8(z, num_frames=3)
    want = ((ref / 2 + 0.5).clamp(0, 1) * 255).round().permute(0, 2, 3, 1)
    assert (u8.float() - want).abs().max() <= 2  # fp16 activations: at most 2 grey levels off
    # single-frame chunk (the reference's last chunk of 25 = 8+8+8+1, quirk Q8)
    with torch.no_grad():
        ref1 = vae.decode(z[:1], num_frames=1).sample
    out1 = nat.decode(z[:1], num_frames=1).sample
    assert ((out1 - ref1).abs().max() / ref1.abs().max()).item() < 1e-2


def test_native_encoder_matches_torch_module():
    """Encoder + quant_conv -> latent mean, incl. the asymmetric-pad stride-2 convs."""
    vae = make_vae(seed=2)
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(1, 3, 64, 96, generator=g) * 2 - 1).half().float()
    with torch.no_grad():
        ref = vae.encode(x).latent_dist.mode()
    nat = NativeTemporalDecoderVAE(vae, ops=ref_ops, device="cpu")
    out = nat.encode(x).latent_dist.mode()
    assert out.shape == ref.shape == (1, 4, 8, 12)
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    assert err < 1e-2, err
