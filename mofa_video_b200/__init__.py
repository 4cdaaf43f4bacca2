"""mofa_video_b200: the MOFA-Video hot path (SVD denoise loop + MOFA-Adapter + CMP + VAE) as hand-written sm_100a kernels
behind a C ABI (include/mofa_b200.h) and the reference's own Python entry points (pipeline/, models/, utils/).
See DESIGN.md and INTEGRATION.md at the repository root."""
