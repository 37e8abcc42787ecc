"""CPU oracle of the baseline-JPEG tile decode (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this; the product path is stamp_b200/jpeg.py + csrc/jpeg_host.cu /
csrc/jpeg.cu).

Restates what ``PIL.Image.open(tile).load()`` does for the cached tiles at the reference's call site
src/stamp/preprocessing/tiling.py:380-406 (tiles were written by ``Image.save(format="jpeg")``: baseline
sequential DCT, Huffman coding, one interleaved scan, 4:2:0 chroma by default).  Pillow decodes through
libjpeg-turbo (third-party, not part of /root/reference; Pillow 12.2.0 bundles libjpeg-turbo 3.x) with the library
defaults, whose algorithm is:
  * entropy decoding per ITU-T T.81 F.2.2 (jdhuff.c);
  * ``jpeg_idct_islow`` (jidctint.c): dequantisation folded into a column pass with 13-bit constants, PASS1_BITS = 2,
    then a row pass, descale by 18 bits, +128, clamp;
  * ``h2v2_fancy_upsample`` (jdsample.c): 3/4 - 1/4 triangle filter, vertical then horizontal, rounding biases
    8 / 7 alternating, first / last column and first / last row replicated (jdmainct.c context rows);
  * ``ycc_rgb_convert`` (jdcolor.c): 16-bit fixed-point tables.
Pinned: tests/test_jpeg_cpu.py compares it bit for bit with Pillow on H&E-like and noise tiles (4:2:0, 4:4:4,
several qualities, restart intervals, sizes that are not multiples of the MCU).
"""

from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,
                   7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
                   39, 46, 53, 60, 61, 54, 47, 55, 62, 63])


class _Bits:
    def __init__(self, data: bytes, pos: int) -> None:
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _fill(self) -> None:
        b = self.d[self.p]
        self.p += 1
        if b == 0xFF:
            nxt = self.d[self.p]
            if nxt == 0:
                self.p += 1
            else:               # a marker inside the entropy-coded segment: feed zeros (T.81 F.2.2.5)
                self.p -= 1
                b = 0
        self.acc = (self.acc << 8) | b
        self.n += 8

    def get(self, k: int) -> int:
        while self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def restart(self) -> None:
        self.acc = self.n = 0
        while not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _huff_table(counts, symbols):
    """code -> symbol lookup as (mincode, maxcode, valptr) per length (T.81 F.2.2.3)."""
    code, k, table = 0, 0, {}
    for length in range(1, 17):
        for _ in range(counts[length - 1]):
            table[(length, code)] = symbols[k]
            code += 1
            k += 1
        code <<= 1
    return table


def _decode_symbol(bits: _Bits, table) -> int:
    code = 0
    for length in range(1, 17):
        code = (code << 1) | bits.get(1)
        if (length, code) in table:
            return table[(length, code)]
    raise ValueError("bad Huffman code")


def _extend(v: int, t: int) -> int:
    return v - (1 << t) + 1 if t and v < (1 << (t - 1)) else v


def parse(data: bytes):
    """-> (height, width, components [(id, h, v, tq)], quant {tq: int32[64] natural order},
           coefficients {comp index: int16 [blocks_y, blocks_x, 64] natural order, quantised})."""
    assert data[:2] == b"\xff\xd8", "not a JPEG"
    p = 2
    quant, dc_tabs, ac_tabs = {}, {}, {}
    restart_interval = 0
    frame = None
    while True:
        assert data[p] == 0xFF
        marker = data[p + 1]
        p += 2
        if marker == 0xD8 or 0xD0 <= marker <= 0xD7 or marker == 0x01:
            continue
        length = int.from_bytes(data[p:p + 2], "big")
        seg = data[p + 2:p + length]
        if marker == 0xDB:
            q = 0
            while q < len(seg):
                pq, tq = seg[q] >> 4, seg[q] & 15
                if pq:
                    vals = np.frombuffer(seg[q + 1:q + 129], dtype=">u2").astype(np.int32)
                    q += 129
                else:
                    vals = np.frombuffer(seg[q + 1:q + 65], dtype=np.uint8).astype(np.int32)
                    q += 65
                nat = np.zeros(64, np.int32)
                nat[ZIGZAG] = vals
                quant[tq] = nat
        elif marker in (0xC0, 0xC1):
            height, width = int.from_bytes(seg[1:3], "big"), int.from_bytes(seg[3:5], "big")
            comps = [(seg[6 + 3 * i], seg[7 + 3 * i] >> 4, seg[7 + 3 * i] & 15, seg[8 + 3 * i]) for i in range(seg[5])]
            frame = (height, width, comps)
        elif marker in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise ValueError("only baseline / extended sequential Huffman JPEG is supported")
        elif marker == 0xC4:
            q = 0
            while q < len(seg):
                tc, th = seg[q] >> 4, seg[q] & 15
                counts = list(seg[q + 1:q + 17])
                n = sum(counts)
                (ac_tabs if tc else dc_tabs)[th] = _huff_table(counts, list(seg[q + 17:q + 17 + n]))
                q += 17 + n
        elif marker == 0xDD:
            restart_interval = int.from_bytes(seg[:2], "big")
        elif marker == 0xDA:
            ns = seg[0]
            sel = {seg[1 + 2 * i]: (seg[2 + 2 * i] >> 4, seg[2 + 2 * i] & 15) for i in range(ns)}
            p += length
            break
        p += length
    height, width, comps = frame
    assert len(sel) == len(comps), "one interleaved scan expected"
    hmax, vmax = max(c[1] for c in comps), max(c[2] for c in comps)
    mcux, mcuy = -(-width // (8 * hmax)), -(-height // (8 * vmax))
    coef = {i: np.zeros((mcuy * c[2], mcux * c[1], 64), np.int16) for i, c in enumerate(comps)}
    bits = _Bits(data, p)
    pred = [0] * len(comps)
    for m in range(mcux * mcuy):
        if restart_interval and m and m % restart_interval == 0:
            bits.restart()
            pred = [0] * len(comps)
        my, mx = divmod(m, mcux)
        for i, (cid, h, v, _tq) in enumerate(comps):
            td, ta = sel[cid]
            for by in range(v):
                for bx in range(h):
                    blk = np.zeros(64, np.int32)
                    t = _decode_symbol(bits, dc_tabs[td])
                    pred[i] += _extend(bits.get(t), t) if t else 0
                    blk[0] = pred[i]
                    k = 1
                    while k < 64:
                        rs = _decode_symbol(bits, ac_tabs[ta])
                        r, s = rs >> 4, rs & 15
                        if s == 0:
                            if r != 15:
                                break
                            k += 16
                            continue
                        k += r
                        blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                        k += 1
                    coef[i][my * v + by, mx * h + bx] = blk
    return height, width, comps, quant, coef


def _c(x: float) -> int:
    return int(x * (1 << 13) + 0.5)


F_0_298, F_0_390, F_0_541, F_0_765, F_0_899, F_1_175 = _c(0.298631336), _c(0.390180644), _c(0.541196100), _c(0.765366865), _c(0.899976223), _c(1.175875602)
F_1_501, F_1_847, F_1_961, F_2_053, F_2_562, F_3_072 = _c(1.501321110), _c(1.847759065), _c(1.961570560), _c(2.053119869), _c(2.562915447), _c(3.072711026)


def _idct_1d(v, shift: int, pre: int):
    """One pass of jpeg_idct_islow over the first axis of int64 array v [8, ...]; ``pre`` = left shift of the even
    part's DC terms (CONST_BITS), ``shift`` = descale."""
    z2, z3 = v[2], v[6]
    z1 = (z2 + z3) * F_0_541
    tmp2 = z1 - z3 * F_1_847
    tmp3 = z1 + z2 * F_0_765
    tmp0 = (v[0] + v[4]) << pre
    tmp1 = (v[0] - v[4]) << pre
    t10, t13, t11, t12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    o0, o1, o2, o3 = v[7], v[5], v[3], v[1]
    z1, z2, z3, z4 = o0 + o3, o1 + o2, o0 + o2, o1 + o3
    z5 = (z3 + z4) * F_1_175
    o0, o1, o2, o3 = o0 * F_0_298, o1 * F_2_053, o2 * F_3_072, o3 * F_1_501
    z1, z2, z3, z4 = -z1 * F_0_899, -z2 * F_2_562, -z3 * F_1_961 + z5, -z4 * F_0_390 + z5
    o0, o1, o2, o3 = o0 + z1 + z3, o1 + z2 + z4, o2 + z2 + z3, o3 + z1 + z4
    rnd = 1 << (shift - 1)
    return np.stack([(t10 + o3 + rnd) >> shift, (t11 + o2 + rnd) >> shift, (t12 + o1 + rnd) >> shift,
                     (t13 + o0 + rnd) >> shift, (t13 - o0 + rnd) >> shift, (t12 - o1 + rnd) >> shift,
                     (t11 - o2 + rnd) >> shift, (t10 - o3 + rnd) >> shift])


def idct_islow(coef: np.ndarray, quant: np.ndarray) -> np.ndarray:
    """int16 [by, bx, 64] quantised coefficients -> uint8 plane [by*8, bx*8]."""
    by, bx, _ = coef.shape
    blk = (coef.astype(np.int64) * quant.astype(np.int64)).reshape(by, bx, 8, 8)     # [.., row, col]
    ws = _idct_1d(np.moveaxis(blk, 2, 0), 13 - 2, 13)                                # columns: over the row index
    ws = np.moveaxis(ws, 0, 2)                                                         # back to [by, bx, row, col]
    out = _idct_1d(np.moveaxis(ws, 3, 0), 13 + 2 + 3, 13)                            # rows: over the column index
    out = np.moveaxis(out, 0, 3)
    out = np.clip(out + 128, 0, 255).astype(np.uint8)
    return out.transpose(0, 2, 1, 3).reshape(by * 8, bx * 8)


def h2v2_fancy_upsample(plane: np.ndarray, ds_h: int, ds_w: int) -> np.ndarray:
    """chroma plane (at least [ds_h, ds_w]) -> [2*ds_h, 2*ds_w]."""
    p = plane[:ds_h, :ds_w].astype(np.int32)
    above = np.vstack([p[:1], p[:-1]])
    below = np.vstack([p[1:], p[-1:]])
    rows = np.empty((2 * ds_h, ds_w), np.int32)
    rows[0::2] = 3 * p + above
    rows[1::2] = 3 * p + below
    last = np.hstack([rows[:, :1], rows[:, :-1]])
    nxt = np.hstack([rows[:, 1:], rows[:, -1:]])
    out = np.empty((2 * ds_h, 2 * ds_w), np.int32)
    out[:, 0::2] = (3 * rows + last + 8) >> 4
    out[:, 1::2] = (3 * rows + nxt + 7) >> 4
    out[:, 0] = (4 * rows[:, 0] + 8) >> 4
    out[:, -1] = (4 * rows[:, -1] + 7) >> 4
    return out.astype(np.uint8)


def _fix(x: float) -> int:
    return int(x * (1 << 16) + 0.5)


def ycc_to_rgb(y: np.ndarray, cb: np.ndarray, cr: np.ndarray) -> np.ndarray:
    y, cb, cr = y.astype(np.int64), cb.astype(np.int64) - 128, cr.astype(np.int64) - 128
    r = y + ((_fix(1.40200) * cr + (1 << 15)) >> 16)
    g = y + ((-_fix(0.34414) * cb + (1 << 15) - _fix(0.71414) * cr) >> 16)
    b = y + ((_fix(1.77200) * cb + (1 << 15)) >> 16)
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def decode(data: bytes) -> np.ndarray:
    """JPEG bytes -> uint8 [H, W, 3] (RGB) or [H, W] (grayscale), as Pillow returns it."""
    height, width, comps, quant, coef = parse(data)
    planes = [idct_islow(coef[i], quant[c[3]]) for i, c in enumerate(comps)]
    if len(comps) == 1:
        return planes[0][:height, :width]
    hmax, vmax = comps[0][1], comps[0][2]
    if (hmax, vmax) == (2, 2) and all(c[1:3] == (1, 1) for c in comps[1:]):
        ds_h, ds_w = -(-height // 2), -(-width // 2)
        cb, cr = (h2v2_fancy_upsample(pl, ds_h, ds_w) for pl in planes[1:])
    elif (hmax, vmax) == (1, 1):
        cb, cr = planes[1], planes[2]
    else:
        raise ValueError("only 4:2:0 and 4:4:4 sampling are restated")
    return ycc_to_rgb(planes[0][:height, :width], cb[:height, :width], cr[:height, :width])
