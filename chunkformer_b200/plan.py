"""Host-side masked-batch packer (thin wrapper over cf_plan_* of the C ABI; pure host code, no GPU needed).

Mirrors the packer / bound tables inside ChunkFormerEncoder.forward_parallel_chunk
(chunkformer/modules/encoder.py:538-612, 627-645) and the padded-batch geometry of forward_encoder
(encoder.py:220-274)."""
import ctypes
from ctypes import POINTER, c_int32, c_int64, c_uint8, c_void_p
from typing import List, Optional, Sequence

import numpy as np

from . import lib as _lib


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(POINTER(c_int32))


class Plan:
    """Chunk list + per-chunk bound tables for one encoder call."""

    def __init__(self, chunk_size: int, left_context: int, right_context: int, lens: Sequence[int],
                 offsets: Optional[Sequence[int]] = None, conv_kernel: int = 15,
                 feat_row_offsets: Optional[Sequence[int]] = None, padded_T: Optional[int] = None):
        L = _lib.load()
        self.c, self.l, self.r, self.kernel = int(chunk_size), int(left_context), int(right_context), int(conv_kernel)
        self.lens = [int(t) for t in lens]
        self.B = len(self.lens)
        self.padded_T = padded_T
        handle = c_void_p()
        lens_a, lens_p = _i32(self.lens)
        if padded_T is None:
            off_p = None
            if offsets is not None:
                off_a, off_p = _i32(list(offsets))
            fro_p = None
            if feat_row_offsets is not None:
                fro_a = np.ascontiguousarray(feat_row_offsets, dtype=np.int64)
                fro_p = fro_a.ctypes.data_as(POINTER(c_int64))
            rc = L.cf_plan_create(self.c, self.l, self.r, self.kernel, self.B, lens_p, off_p, fro_p, ctypes.byref(handle))
        else:
            rc = L.cf_plan_create_padded(self.c, self.l, self.r, self.kernel, self.B, int(padded_T), lens_p,
                                         ctypes.byref(handle))
        _lib.check(rc, None, "cf_plan_create")
        self._h = handle
        self.n = L.cf_plan_num_chunks(self._h)
        self.rows = L.cf_plan_rows(self._h)
        nck = np.zeros(self.B, dtype=np.int32)
        el = np.zeros(self.B, dtype=np.int32)
        _lib.check(L.cf_plan_tables(self._h, nck.ctypes.data_as(POINTER(c_int32)), el.ctypes.data_as(POINTER(c_int32))))
        self.n_chunks: List[int] = [int(v) for v in nck]
        self.enc_lens = el

    @property
    def handle(self):
        return self._h

    def masks(self):
        """(att_mask (n, l+c+r) bool, conv_mask (n, c+2*lorder) bool) as the reference hands them to the layers."""
        L = _lib.load()
        W, CW = self.l + self.c + self.r, self.c + 2 * (self.kernel // 2)
        att = np.zeros((self.n, W), dtype=np.uint8)
        cv = np.zeros((self.n, CW), dtype=np.uint8)
        _lib.check(L.cf_plan_masks(self._h, att.ctypes.data_as(POINTER(c_uint8)), cv.ctypes.data_as(POINTER(c_uint8))))
        return att.astype(bool), cv.astype(bool)

    def chunk_table(self) -> np.ndarray:
        """(n, 8) int32: utt, j, att_lo, att_hi, conv_lo, conv_hi, out_lo, out_hi."""
        t = np.zeros((self.n, 8), dtype=np.int32)
        _lib.check(_lib.load().cf_plan_chunk_table(self._h, t.ctypes.data_as(POINTER(c_int32))))
        return t

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().cf_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass
