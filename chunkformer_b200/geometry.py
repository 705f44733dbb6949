"""Encoder geometry of the hot path (ChunkFormerEncoder ctor arguments that matter at inference).

Mirrors the constructor of the reference encoder (chunkformer/modules/encoder.py:36-193) and the
`encoder_conf` block of its YAML configs (e.g. examples/asr/rnnt/conf/chunkformer-rnnt-large-vie.yaml:4-24).
"""
from dataclasses import dataclass, asdict

SUBSAMPLING = 8          # dw_striding, subsampling.py:43-45
SUB_CONTEXT = 15         # right_context + 1, encoder.py:540
FEAT_FREQ = (39, 19, 9)  # 80 -> 39 -> 19 -> 9 (three 3x3 stride-2 valid convs)


@dataclass(frozen=True)
class EncoderGeometry:
    d_model: int = 512
    heads: int = 8
    ffn: int = 2048
    layers: int = 17
    kernel: int = 15
    vocab: int = 5000
    feat_dim: int = 80
    has_cmvn: bool = False
    conv_norm: str = "layer_norm"      # cnn_module_norm: "layer_norm" | "batch_norm" (eval-mode BatchNorm1d, convolution.py:83-89)

    @property
    def d_k(self) -> int:
        return self.d_model // self.heads

    @property
    def lorder(self) -> int:
        return self.kernel // 2

    def to_dict(self):
        return asdict(self)

    @classmethod
    def from_encoder_conf(cls, encoder_conf: dict, input_dim: int = 80, vocab: int = 0,
                          has_cmvn: bool = False) -> "EncoderGeometry":
        """Build from a reference `encoder_conf`; reject what the reference's masked-batch path
        itself cannot run (encoder.py:503-681 needs dynamic_conv) and the
        constructor switches no shipped config uses."""
        ec = dict(encoder_conf)
        if ec.get("input_layer", "dw_striding") != "dw_striding":
            raise ValueError("only input_layer=dw_striding is supported")
        if ec.get("pos_enc_layer_type", "chunk_rel_pos") != "chunk_rel_pos":
            raise ValueError("only pos_enc_layer_type=chunk_rel_pos is supported")
        if ec.get("selfattention_layer_type", "chunk_rel_seflattn") != "chunk_rel_seflattn":
            raise ValueError("only selfattention_layer_type=chunk_rel_seflattn is supported")
        conv_norm = ec.get("cnn_module_norm", "batch_norm")        # the reference constructor's default (encoder.py:59)
        if conv_norm not in ("layer_norm", "batch_norm"):
            raise ValueError("cnn_module_norm must be layer_norm or batch_norm")
        if not ec.get("dynamic_conv", False):
            raise ValueError("dynamic_conv must be true (the reference's masked-batch path requires it)")
        if not ec.get("normalize_before", True) or not ec.get("macaron_style", True):
            raise ValueError("normalize_before and macaron_style must be true")
        if not ec.get("use_cnn_module", True) or ec.get("causal", False):
            raise ValueError("use_cnn_module must be true and causal false")
        if ec.get("layer_norm_type", "layer_norm") != "layer_norm":
            raise ValueError("only layer_norm_type=layer_norm is supported")
        if ec.get("activation_type", "swish") != "swish":
            raise ValueError("only activation_type=swish is supported")
        return cls(d_model=ec.get("output_size", 256), heads=ec.get("attention_heads", 4),
                   ffn=ec.get("linear_units", 2048), layers=ec.get("num_blocks", 6),
                   kernel=ec.get("cnn_module_kernel", 15), vocab=vocab, feat_dim=input_dim,
                   has_cmvn=has_cmvn, conv_norm=conv_norm)


CTC_LARGE = EncoderGeometry(512, 8, 2048, 17, 15, 5000)       # docs/paper.pdf III.A, Table III
RNNT_LARGE = EncoderGeometry(512, 4, 2048, 12, 15, 1024)      # chunkformer-rnnt-large-vie.yaml:4-24
CTC_SMALL = EncoderGeometry(256, 4, 2048, 12, 15, 5000)       # chunkformer-ctc-small-libri-960h.yaml:4-24
