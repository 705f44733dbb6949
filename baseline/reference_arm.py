"""The UNMODIFIED reference (ishine/chunkformer) as a baseline arm.

`baseline/_ref/` holds the reference installed once in the build container with
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref <copy of /root/reference>
(git-ignored, not gpurun-ignored: it travels to the GPU box; /root/reference itself does not exist there).  This module only
imports it and drives its own public code path: `ChunkFormerEncoder.forward_parallel_chunk` (modules/encoder.py:503-681)
followed by `ctc.log_softmax(...).argmax` (modules/ctc.py:73-91; call sites chunkformer_model.py:437-438, 526-527), on random
weights in the reference's checkpoint layout (chunkformer_b200.synth.synth_state_dict).  None of this repo's kernels, models or
oracle are on that path.

The reference's package __init__ imports `jiwer`, `colorama` and `pydub` (CLI / audio-file loading only); they are absent in
this image and unused on the encoder path, so three empty stub modules are registered before the import.
"""
import os
import sys
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available(root: str = REF_ROOT) -> bool:
    return os.path.isdir(os.path.join(root, "chunkformer"))


def import_reference(root: str = REF_ROOT):
    if not available(root):
        raise RuntimeError(f"reference not installed under {root}")
    for name in ("jiwer", "colorama", "pydub"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["colorama"].Fore = types.SimpleNamespace(RED="", GREEN="", YELLOW="")
    sys.modules["colorama"].Style = types.SimpleNamespace(RESET_ALL="")
    sys.modules["pydub"].AudioSegment = object
    if root not in sys.path:
        sys.path.insert(0, root)
    from chunkformer import ChunkFormerModel  # noqa: E402
    from chunkformer.chunkformer_model import ChunkFormerConfig  # noqa: E402
    return ChunkFormerModel, ChunkFormerConfig


def reference_config_dict(d=512, heads=8, ffn=2048, layers=17, vocab=5000, kernel=15, conv_norm="layer_norm"):
    """Config dict for a random-init reference ASR model; the attention decoder is unused on the path
    and kept minimal."""
    return dict(
        input_dim=80, output_dim=vocab, model="asr_model", encoder="chunkformer",
        encoder_conf=dict(
            output_size=d, attention_heads=heads, linear_units=ffn, num_blocks=layers,
            dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1,
            input_layer="dw_striding", normalize_before=True, cnn_module_kernel=kernel,
            use_cnn_module=True, activation_type="swish", pos_enc_layer_type="chunk_rel_pos",
            selfattention_layer_type="chunk_rel_seflattn", cnn_module_norm=conv_norm,
            dynamic_conv=True),
        decoder="bitransformer",
        decoder_conf=dict(attention_heads=4, linear_units=64, num_blocks=1, r_num_blocks=1,
                          dropout_rate=0.1, positional_dropout_rate=0.1,
                          self_attention_dropout_rate=0.1, src_attention_dropout_rate=0.1),
        ctc="ctc", ctc_conf=dict(ctc_blank_id=0),
        model_conf=dict(ctc_weight=0.3, lsm_weight=0.1, length_normalized_loss=False,
                        reverse_weight=0.3))


def build_reference(geo, state_dict, root: str = REF_ROOT):
    """Reference model of geometry `geo` carrying `state_dict` (encoder.* / ctc.* keys; the unused decoder stays random)."""
    Model, Config = import_reference(root)
    cfg = reference_config_dict(geo.d_model, geo.heads, geo.ffn, geo.layers, geo.vocab, geo.kernel, geo.conv_norm)
    model = Model(Config.from_dict(cfg)).eval()
    if geo.has_cmvn:
        from chunkformer.modules.cmvn import GlobalCMVN
        model.model.encoder.global_cmvn = GlobalCMVN(state_dict["encoder.global_cmvn.mean"].clone(),
                                                     state_dict["encoder.global_cmvn.istd"].clone())
    missing, unexpected = model.model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("decoder.") for k in missing), [k for k in missing if not k.startswith("decoder.")]
    return model


@torch.no_grad()
def reference_step(model, xs, lens, c, l, r, device="cpu", autocast_dtype=None):
    """One pass of the reference's own hot path over one masked batch: forward_parallel_chunk + greedy CTC; returns the
    flat (n, c) token ids on `device`."""
    enc, ctc = model.model.encoder, model.model.ctc
    lens_t = torch.tensor(lens, dtype=torch.int, device=device)
    offset = torch.zeros(len(lens), dtype=torch.int, device=device)
    ctx = torch.autocast(torch.device(device).type, dtype=autocast_dtype) if autocast_dtype is not None else _Null()
    with ctx:                              # chunkformer_model.py:708-743 wraps its decode loop the same way
        out, enc_lens, n_chunks, _, _, _ = enc.forward_parallel_chunk(
            xs=xs, xs_origin_lens=lens_t, chunk_size=c, left_context_size=l, right_context_size=r, offset=offset)
        tokens = ctc.log_softmax(out).argmax(-1)
    return tokens, enc_lens, n_chunks


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def time_reference_gpu(model, xs_dev, lens, c, l, r, device, autocast_dtype, steps=2, warmup=1):
    """CUDA-event time of the reference's eager GPU path for one masked batch (ms per step, peak memory in bytes)."""
    torch.cuda.reset_peak_memory_stats(device)
    for _ in range(warmup):
        reference_step(model, xs_dev, lens, c, l, r, device, autocast_dtype)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tok, _, _ = reference_step(model, xs_dev, lens, c, l, r, device, autocast_dtype)
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / steps, int(torch.cuda.max_memory_allocated(device)), tok


def time_reference_cpu(model, xs, lens, c, l, r, steps, warmup):
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        reference_step(model, xs, lens, c, l, r, "cpu", None)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times
