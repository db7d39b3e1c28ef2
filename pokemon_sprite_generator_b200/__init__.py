"""B200-native latent-diffusion hot path (U-Net denoiser, cosine noising, DDPM step) for
GabrieleConte/pokemon-sprite-generator -- drop-in for `src.models.UNet` and `src.training.DiffusionTrainer`.

Host code is PyTorch (device memory, streams, torch.distributed); all compute is hand-written sm_100a CUDA
behind the C ABI declared in include/psg_b200.h (csrc/libpsg_b200.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"
