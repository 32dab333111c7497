"""numpy restatement of the reference's m16n8k16 layout ops.  TEST INFRASTRUCTURE ONLY.

Lane geometry shared by every op (reference: TinyGemmConvertA.cu:121-136,248-281;
TinyGemmConvertB.cu:48-59,276-303): for lane t of a warp, g = t // 4, q = t % 4,
k0 = kTile*16 + 2*q.

  A fragment (16 rows x 16 k), values v0..v7:
      (g,k0) (g,k0+1) (g+8,k0) (g+8,k0+1) (g,k0+8) (g,k0+9) (g+8,k0+8) (g+8,k0+9)
  B fragment (8 rows x 16 k), values v0..v3:
      (g,k0) (g,k0+1) (g,k0+8) (g,k0+9)

  4-bit pack of v0..v7 : v7<<28 | v5<<24 | v3<<20 | v1<<16 | v6<<12 | v4<<8 | v2<<4 | v0
  8-bit pack of v0..v3 : v3<<24 | v1<<16 | v2<<8 | v0

Out-of-range rows / columns read as 0.  16-bit payloads are handled as uint16 bit
patterns so that the same code serves bf16 and fp16 bit-exactly.
"""
import numpy as np

_T = np.arange(32)
_G = _T // 4
_Q = _T % 4

# (row offset, k offset) of v0..v7 in an A fragment, minus the lane terms
_A_ROW = np.array([0, 0, 8, 8, 0, 0, 8, 8])
_A_COL = np.array([0, 1, 0, 1, 8, 9, 8, 9])
# k offsets of v0..v3 in a B fragment
_B_COL = np.array([0, 1, 8, 9])

_NIB_SHIFT = np.array([0, 16, 4, 20, 8, 24, 12, 28], dtype=np.uint32)  # shift of v_i
_BYTE_SHIFT = np.array([0, 16, 8, 24], dtype=np.uint32)


def div_up(a, b):
    return (a + b - 1) // b


def _gather2d(src, rows, cols):
    """src[rows, cols] with zeros where out of range (any integer/uint dtype)."""
    m, k = src.shape
    ok = (rows < m) & (cols < k)
    r = np.where(ok, rows, 0)
    c = np.where(ok, cols, 0)
    out = src[r, c]
    return np.where(ok, out, np.zeros((), dtype=src.dtype))


# ----------------------------------------------------------------------------
# 16-bit A layout   [m][k] <-> [ceil(m/16)][ceil(k/16)][32][8]
# reference: TinyGemmConvertA.cu:19-141 (to), :442-546 (from)
# ----------------------------------------------------------------------------
def to_A(x_u16, inner_k_tiles=1):
    assert inner_k_tiles == 1
    m, k = x_u16.shape
    mT, kT = div_up(m, 16), div_up(k, 16)
    mt = np.arange(mT)[:, None, None, None]
    kt = np.arange(kT)[None, :, None, None]
    g = _G[None, None, :, None]
    q = _Q[None, None, :, None]
    rows = mt * 16 + g + _A_ROW[None, None, None, :]
    cols = kt * 16 + 2 * q + _A_COL[None, None, None, :]
    return _gather2d(x_u16, rows, cols)


def from_A(a_u16, m, k):
    mT, kT = a_u16.shape[:2]
    assert (mT, kT) == (div_up(m, 16), div_up(k, 16)) and a_u16.shape[2:] == (32, 8)
    out = np.zeros((mT * 16, kT * 16), dtype=a_u16.dtype)
    mt = np.arange(mT)[:, None, None, None]
    kt = np.arange(kT)[None, :, None, None]
    g = _G[None, None, :, None]
    q = _Q[None, None, :, None]
    rows = np.broadcast_to(mt * 16 + g + _A_ROW[None, None, None, :], a_u16.shape)
    cols = np.broadcast_to(kt * 16 + 2 * q + _A_COL[None, None, None, :], a_u16.shape)
    out[rows, cols] = a_u16
    return out[:m, :k].copy()


# ----------------------------------------------------------------------------
# 16-bit B layout   [n][k] <-> [ceil(n/8)][ceil(k/(ik*16))][32][ik*4]
# reference: TinyGemmConvertB.cu:20-66 (to), :136-176 (from)
# ----------------------------------------------------------------------------
def to_B(x_u16, inner_k_tiles):
    assert inner_k_tiles in (1, 2)
    ik = inner_k_tiles
    n, k = x_u16.shape
    nT, kO = div_up(n, 8), div_up(k, 16 * ik)
    nt = np.arange(nT)[:, None, None, None, None]
    ko = np.arange(kO)[None, :, None, None, None]
    g = _G[None, None, :, None, None]
    q = _Q[None, None, :, None, None]
    ki = np.arange(ik)[None, None, None, :, None]
    rows = nt * 8 + g + 0 * ki
    cols = (ko * ik + ki) * 16 + 2 * q + _B_COL[None, None, None, None, :]
    out = _gather2d(x_u16, np.broadcast_to(rows, np.broadcast_shapes(rows.shape, cols.shape)), cols)
    return out.reshape(nT, kO, 32, ik * 4)


def from_B(b_u16, n, k):
    nT, kO, _, last = b_u16.shape
    ik = last // 4
    assert ik in (1, 2) and last % 4 == 0 and b_u16.shape[2] == 32
    assert nT == div_up(n, 8) and kO == div_up(k, 16 * ik)
    v = b_u16.reshape(nT, kO, 32, ik, 4)
    out = np.zeros((nT * 8, kO * ik * 16), dtype=b_u16.dtype)
    nt = np.arange(nT)[:, None, None, None, None]
    ko = np.arange(kO)[None, :, None, None, None]
    g = _G[None, None, :, None, None]
    q = _Q[None, None, :, None, None]
    ki = np.arange(ik)[None, None, None, :, None]
    rows = np.broadcast_to(nt * 8 + g + 0 * ki + 0 * _B_COL, v.shape)
    cols = np.broadcast_to((ko * ik + ki) * 16 + 2 * q + _B_COL, v.shape)
    out[rows, cols] = v
    return out[:n, :k].copy()


# ----------------------------------------------------------------------------
# packed 4-bit A layout  [m][k] int32 codes -> [ceil(m/16)][ceil(k/(ik*16))][32][ik] int32
# reference: TinyGemmConvertA.cu:226-285
# ----------------------------------------------------------------------------
def _a_frag_codes(codes, mT, kT_total):
    """[mT][kT_total][32][8] uint32 gather of A-fragment values (zero padded)."""
    src = codes.astype(np.int64).astype(np.uint32)  # int32 -> uint32 wrap like the C cast
    mt = np.arange(mT)[:, None, None, None]
    kt = np.arange(kT_total)[None, :, None, None]
    g = _G[None, None, :, None]
    q = _Q[None, None, :, None]
    rows = mt * 16 + g + _A_ROW[None, None, None, :]
    cols = kt * 16 + 2 * q + _A_COL[None, None, None, :]
    return _gather2d(src, rows, cols)


def _b_frag_codes(codes, nT, kT_total):
    """[nT][kT_total][32][4] uint32 gather of B-fragment values (zero padded)."""
    src = codes.astype(np.int64).astype(np.uint32)
    nt = np.arange(nT)[:, None, None, None]
    kt = np.arange(kT_total)[None, :, None, None]
    g = _G[None, None, :, None]
    q = _Q[None, None, :, None]
    rows = nt * 8 + g + 0 * _B_COL[None, None, None, :]
    cols = kt * 16 + 2 * q + _B_COL[None, None, None, :]
    return _gather2d(src, np.broadcast_to(rows, np.broadcast_shapes(rows.shape, cols.shape)), cols)


def _pack_nibbles(v8):
    """v8[..., 8] uint32 -> uint32 word, exactly (v<<s) OR-ed with uint32 wrap."""
    w = np.zeros(v8.shape[:-1], dtype=np.uint32)
    for i in range(8):
        w |= (v8[..., i] << _NIB_SHIFT[i]).astype(np.uint32)
    return w


def _pack_bytes(v4):
    w = np.zeros(v4.shape[:-1], dtype=np.uint32)
    for i in range(4):
        w |= (v4[..., i] << _BYTE_SHIFT[i]).astype(np.uint32)
    return w


def to_Aint4(codes, inner_k_tiles):
    ik = inner_k_tiles
    assert ik in (1, 2, 4)
    m, k = codes.shape
    mT, kS = div_up(m, 16), div_up(k, 16 * ik)
    frag = _a_frag_codes(codes, mT, kS * ik)                # [mT][kS*ik][32][8]
    words = _pack_nibbles(frag)                             # [mT][kS*ik][32]
    out = words.reshape(mT, kS, ik, 32).transpose(0, 1, 3, 2)
    return np.ascontiguousarray(out).view(np.int32)


# reference: TinyGemmConvertA.cu:340-396.  NOTE the reference launches one block per
# *valid* k-tile, so when ceil(k/16) is not a multiple of ik the trailing inner slots
# of the last outer tile are left uninitialised by the reference; here they are 0.
def to_Aint8(codes, inner_k_tiles):
    ik = inner_k_tiles
    assert ik in (1, 2)
    m, k = codes.shape
    mT, kT = div_up(m, 16), div_up(k, 16)
    kO = div_up(kT, ik)
    frag = _a_frag_codes(codes, mT, kO * ik)                # [mT][kO*ik][32][8]
    w0 = _pack_bytes(frag[..., 0:4])
    w1 = _pack_bytes(frag[..., 4:8])
    words = np.stack([w0, w1], axis=-1)                     # [mT][kO*ik][32][2]
    out = words.reshape(mT, kO, ik, 32, 2).transpose(0, 1, 3, 2, 4).reshape(mT, kO, 32, ik * 2)
    return np.ascontiguousarray(out).view(np.int32)


# reference: TinyGemmConvertB.cu:252-308 (requires k % (ik*16) == 0, :337)
def to_Bint4(codes, inner_k_tiles):
    ik = inner_k_tiles
    assert ik in (2, 4, 8)
    n, k = codes.shape
    assert k % (ik * 16) == 0
    nT, kS = div_up(n, 8), k // (ik * 16)
    frag = _b_frag_codes(codes, nT, kS * ik)                # [nT][kS*ik][32][4]
    pair = frag.reshape(nT, kS, ik // 2, 2, 32, 4).transpose(0, 1, 2, 4, 3, 5)
    v8 = pair.reshape(nT, kS, ik // 2, 32, 8)               # tile 2j (v0-3) || tile 2j+1 (v4-7)
    words = _pack_nibbles(v8)                               # [nT][kS][ik/2][32]
    out = words.transpose(0, 1, 3, 2)
    return np.ascontiguousarray(out).view(np.int32)


# reference: TinyGemmConvertB.cu:367-410 (requires k % (ik*16) == 0, :441)
def to_Bint8(codes, inner_k_tiles):
    ik = inner_k_tiles
    assert ik in (1, 2, 4)
    n, k = codes.shape
    assert k % (ik * 16) == 0
    nT, kS = div_up(n, 8), k // (ik * 16)
    frag = _b_frag_codes(codes, nT, kS * ik)                # [nT][kS*ik][32][4]
    words = _pack_bytes(frag)                               # [nT][kS*ik][32]
    out = words.reshape(nT, kS, ik, 32).transpose(0, 1, 3, 2)
    return np.ascontiguousarray(out).view(np.int32)


# ----------------------------------------------------------------------------
# Inverses of the packed layouts (the reference has none; the GEMM kernels consume the
# packed words directly).  Used by the oracle GEMM to recover row-major codes.
# ----------------------------------------------------------------------------
def _unpack_nibbles(words_u32):
    return np.stack([(words_u32 >> s) & np.uint32(0xF) for s in _NIB_SHIFT], axis=-1)


def _unpack_bytes(words_u32):
    return np.stack([(words_u32 >> s) & np.uint32(0xFF) for s in _BYTE_SHIFT], axis=-1)


def from_Aint4(packed):
    """[mT][kS][32][ik] int32 -> codes [mT*16][kS*ik*16] int32."""
    mT, kS, _, ik = packed.shape
    w = np.ascontiguousarray(packed).view(np.uint32).transpose(0, 1, 3, 2).reshape(mT, kS * ik, 32)
    v8 = _unpack_nibbles(w)                                  # [mT][kT][32][8]
    out = np.zeros((mT * 16, kS * ik * 16), dtype=np.int32)
    mt = np.arange(mT)[:, None, None, None]
    kt = np.arange(kS * ik)[None, :, None, None]
    g = _G[None, None, :, None]
    q = _Q[None, None, :, None]
    rows = np.broadcast_to(mt * 16 + g + _A_ROW, v8.shape)
    cols = np.broadcast_to(kt * 16 + 2 * q + _A_COL, v8.shape)
    out[rows, cols] = v8.astype(np.int32)
    return out


def from_Aint8(packed):
    mT, kO, _, last = packed.shape
    ik = last // 2
    w = np.ascontiguousarray(packed).view(np.uint32).reshape(mT, kO, 32, ik, 2)
    w = w.transpose(0, 1, 3, 2, 4).reshape(mT, kO * ik, 32, 2)
    v8 = np.concatenate([_unpack_bytes(w[..., 0]), _unpack_bytes(w[..., 1])], axis=-1)
    out = np.zeros((mT * 16, kO * ik * 16), dtype=np.int32)
    mt = np.arange(mT)[:, None, None, None]
    kt = np.arange(kO * ik)[None, :, None, None]
    g = _G[None, None, :, None]
    q = _Q[None, None, :, None]
    rows = np.broadcast_to(mt * 16 + g + _A_ROW, v8.shape)
    cols = np.broadcast_to(kt * 16 + 2 * q + _A_COL, v8.shape)
    out[rows, cols] = v8.astype(np.int32)
    return out


def from_Bint4(packed):
    """[nT][kS][32][ik/2] int32 -> codes [nT*8][kS*ik*16] int32."""
    nT, kS, _, half = packed.shape
    ik = half * 2
    w = np.ascontiguousarray(packed).view(np.uint32).transpose(0, 1, 3, 2)   # [nT][kS][ik/2][32]
    v8 = _unpack_nibbles(w)                                                   # [..][32][8]
    v = v8.reshape(nT, kS, half, 32, 2, 4).transpose(0, 1, 2, 4, 3, 5).reshape(nT, kS * ik, 32, 4)
    out = np.zeros((nT * 8, kS * ik * 16), dtype=np.int32)
    nt = np.arange(nT)[:, None, None, None]
    kt = np.arange(kS * ik)[None, :, None, None]
    g = _G[None, None, :, None]
    q = _Q[None, None, :, None]
    rows = np.broadcast_to(nt * 8 + g + 0 * _B_COL, v.shape)
    cols = np.broadcast_to(kt * 16 + 2 * q + _B_COL, v.shape)
    out[rows, cols] = v.astype(np.int32)
    return out


def from_Bint8(packed):
    nT, kS, _, ik = packed.shape
    w = np.ascontiguousarray(packed).view(np.uint32).transpose(0, 1, 3, 2).reshape(nT, kS * ik, 32)
    v = _unpack_bytes(w)
    out = np.zeros((nT * 8, kS * ik * 16), dtype=np.int32)
    nt = np.arange(nT)[:, None, None, None]
    kt = np.arange(kS * ik)[None, :, None, None]
    g = _G[None, None, :, None]
    q = _Q[None, None, :, None]
    rows = np.broadcast_to(nt * 8 + g + 0 * _B_COL, v.shape)
    cols = np.broadcast_to(kt * 16 + 2 * q + _B_COL, v.shape)
    out[rows, cols] = v.astype(np.int32)
    return out
