"""bisinger_b200 -- B200-native (sm_100a) synthesis hot path for BiSinger: shallow-diffusion sampler over DiffNet, the
PitchExtractor (mel -> f0) and the HiFi-GAN/NSF generator, behind the reference's own module signatures.  All compute lives in
``libbisinger_b200.so`` (hand-written CUDA: tcgen05 / TMEM / TMA); see include/bisinger_b200.h and DESIGN.md."""
from ._lib import LIB_PATH, launch_count, lib  # noqa: F401
from .diffusion import B200DiffNet, B200GaussianDiffusion, DiffusionPlan  # noqa: F401
from .infer import mel_to_wav, synthesize  # noqa: F401
from .fft import B200FastspeechDecoder, B200FastspeechEncoder, B200FFTBlocks, device_blocks_encoder  # noqa: F401
from .pitch import B200PitchExtractor  # noqa: F401

__all__ = ["B200DiffNet", "B200GaussianDiffusion", "DiffusionPlan", "B200PitchExtractor", "B200FastspeechDecoder", "B200FastspeechEncoder", "B200FFTBlocks", "device_blocks_encoder", "mel_to_wav", "synthesize", "lib",
           "launch_count", "LIB_PATH"]
