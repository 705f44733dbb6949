"""Import the UNMODIFIED reference (ishine/chunkformer) from /root/reference on CPU.

Only used by tests/golden/make_golden.py (fixture generation in the build container).
Nothing that runs on the GPU box imports this module: /root/reference does not exist there.

The reference's package __init__ pulls `jiwer`, `colorama` and `pydub` (CLI / audio-file
loading only); they are absent in this image and never used on the encoder path, so three
empty stub modules are registered before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CHUNKFORMER_REFERENCE", "/root/reference")


def import_reference():
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "chunkformer")):
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in ("jiwer", "colorama", "pydub"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["colorama"].Fore = types.SimpleNamespace(RED="", GREEN="", YELLOW="")
    sys.modules["colorama"].Style = types.SimpleNamespace(RESET_ALL="")
    sys.modules["pydub"].AudioSegment = object
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from chunkformer import ChunkFormerModel  # noqa: E402
    from chunkformer.chunkformer_model import ChunkFormerConfig  # noqa: E402

    return ChunkFormerModel, ChunkFormerConfig


def reference_config_dict(d=512, heads=8, ffn=2048, layers=17, vocab=5000, kernel=15):
    """Config dict for a random-init reference ASR model."""
    return dict(
        input_dim=80, output_dim=vocab, model="asr_model", encoder="chunkformer",
        encoder_conf=dict(
            output_size=d, attention_heads=heads, linear_units=ffn, num_blocks=layers,
            dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1,
            input_layer="dw_striding", normalize_before=True, cnn_module_kernel=kernel,
            use_cnn_module=True, activation_type="swish", pos_enc_layer_type="chunk_rel_pos",
            selfattention_layer_type="chunk_rel_seflattn", cnn_module_norm="layer_norm",
            dynamic_conv=True),
        decoder="bitransformer",
        decoder_conf=dict(attention_heads=4, linear_units=64, num_blocks=1, r_num_blocks=1,
                          dropout_rate=0.1, positional_dropout_rate=0.1,
                          self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1),
        ctc="ctc", ctc_conf=dict(ctc_blank_id=0),
        model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False,
                        reverse_weight=0.3))
