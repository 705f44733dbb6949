"""chunkformer_b200: B200-native (sm_100a) implementation of ChunkFormer's masked-chunk
Conformer encoder forward + greedy CTC, drop-in behind the reference's Python API."""
from .geometry import EncoderGeometry, CTC_LARGE, RNNT_LARGE, CTC_SMALL  # noqa: F401

__version__ = "0.1.0"
