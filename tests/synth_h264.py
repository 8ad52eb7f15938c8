"""Synthetic H.264 Baseline bitstream writer -- TEST INFRASTRUCTURE.

The reference repository ships three encoder-made streams (test/*.h264): one slice per picture, no FMO/ASO,
no I_PCM, one reference frame.  This module writes small, *valid* Annex-B streams that walk through the rest
of the Baseline syntax the reference decoder accepts (ITU-T H.264 03/2005 clause numbers in the comments):

  * pictures of 1x1 .. 11x9 macroblocks (level 1 .. 3), cropping, VUI incl. HRD and bitstream restriction
  * several slices per picture, arbitrary slice order, slice groups of map type 0..6 (FMO)
  * I / P slices mixed in a picture; every macroblock type: P_Skip, P_L0_16x16/16x8/8x16, P_8x8 with all four
    sub-macroblock types, P_8x8ref0, Intra4x4 (all 9 modes where their neighbours exist), Intra16x16 (4 modes
    x coded patterns), I_PCM; constrained_intra_pred
  * several reference frames, ref_idx per partition, reference list reordering, non-reference pictures,
    long-term IDR, sliding-window marking, frame_num wrap-around, POC types 0, 1, 2 (incl. output reordering)
  * CAVLC residuals incl. level escape codes; coefficient magnitudes bounded so the inverse transform stays
    inside the [-512, 511] check of h264bsdProcessBlock (h264bsd_transform.c:183-188)
  * motion vectors up to the limits the reference accepts (hor [-8192, 8191], ver [-2048, 2047] quarter
    samples, h264bsd_inter_prediction.c:537-545), i.e. far outside the picture
  * mb_qp_delta incl. wrap-around, chroma_qp_index_offset, all three disable_deblocking_filter_idc values
    with filter offsets

It carries its own model of the syntax-level state (neighbour availability, nC, Intra4x4 mode prediction,
motion vector prediction, reference list) because the *encoder* needs it to emit decodable codes; nothing is
shared with the product's host parser or with oracle/px_oracle.c.  The streams carry no meaningful image:
the tests compare the repo's decoder with the compiled reference (oracle/_ref) on them, byte for byte.

    make_stream(seed) -> bytes        deterministic in `seed`
"""
import random

# ---------------------------------------------------------------------------------------------- tables
# Table 9-5 coeff_token (length, code) [vlcTable][trailingOnes][totalCoeff]
kTokLen = [
    [[1, 6, 8, 9, 10, 11, 13, 13, 13, 14, 14, 15, 15, 16, 16, 16, 16],
     [0, 2, 6, 8, 9, 10, 11, 13, 13, 14, 14, 15, 15, 15, 16, 16, 16],
     [0, 0, 3, 7, 8, 9, 10, 11, 13, 13, 14, 14, 15, 15, 16, 16, 16],
     [0, 0, 0, 5, 6, 7, 8, 9, 10, 11, 13, 14, 14, 15, 15, 16, 16]],
    [[2, 6, 6, 7, 8, 8, 9, 11, 11, 12, 12, 12, 13, 13, 13, 14, 14],
     [0, 2, 5, 6, 6, 7, 8, 9, 11, 11, 12, 12, 13, 13, 14, 14, 14],
     [0, 0, 3, 6, 6, 7, 8, 9, 11, 11, 12, 12, 13, 13, 13, 14, 14],
     [0, 0, 0, 4, 4, 5, 6, 6, 7, 9, 11, 11, 12, 13, 13, 13, 14]],
    [[4, 6, 6, 6, 7, 7, 7, 7, 8, 8, 9, 9, 9, 10, 10, 10, 10],
     [0, 4, 5, 5, 5, 5, 6, 6, 7, 8, 8, 9, 9, 9, 10, 10, 10],
     [0, 0, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10],
     [0, 0, 0, 4, 4, 4, 4, 4, 5, 6, 7, 8, 8, 9, 10, 10, 10]]]
kTokCode = [
    [[1, 5, 7, 7, 7, 7, 15, 11, 8, 15, 11, 15, 11, 15, 11, 7, 4],
     [0, 1, 4, 6, 6, 6, 6, 14, 10, 14, 10, 14, 10, 1, 14, 10, 6],
     [0, 0, 1, 5, 5, 5, 5, 5, 13, 9, 13, 9, 13, 9, 13, 9, 5],
     [0, 0, 0, 3, 3, 4, 4, 4, 4, 4, 12, 12, 8, 12, 8, 12, 8]],
    [[3, 11, 7, 7, 7, 4, 7, 15, 11, 15, 11, 8, 15, 11, 7, 9, 7],
     [0, 2, 7, 10, 6, 6, 6, 6, 14, 10, 14, 10, 14, 10, 11, 8, 6],
     [0, 0, 3, 9, 5, 5, 5, 5, 13, 9, 13, 9, 13, 9, 6, 10, 5],
     [0, 0, 0, 5, 4, 6, 8, 4, 4, 4, 12, 8, 12, 12, 8, 1, 4]],
    [[15, 15, 11, 8, 15, 11, 9, 8, 15, 11, 15, 11, 8, 13, 9, 5, 1],
     [0, 14, 15, 12, 10, 8, 14, 10, 14, 14, 10, 14, 10, 7, 12, 8, 4],
     [0, 0, 13, 14, 11, 9, 13, 9, 13, 10, 13, 9, 13, 9, 11, 7, 3],
     [0, 0, 0, 12, 11, 10, 9, 8, 13, 12, 12, 12, 8, 12, 10, 6, 2]]]
kTokDcLen = [[2, 6, 6, 6, 6], [0, 1, 6, 7, 8], [0, 0, 3, 7, 8], [0, 0, 0, 6, 7]]
kTokDcCode = [[1, 7, 4, 3, 2], [0, 1, 6, 3, 3], [0, 0, 1, 2, 2], [0, 0, 0, 5, 0]]
# Tables 9-7 / 9-8 total_zeros for 4x4 blocks [totalCoeff-1][total_zeros]
kTzLen = [
    [1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 9],
    [3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 6],
    [4, 3, 3, 3, 4, 4, 3, 3, 4, 5, 5, 6, 5, 6],
    [5, 3, 4, 4, 3, 3, 3, 4, 3, 4, 5, 5, 5],
    [4, 4, 4, 3, 3, 3, 3, 3, 4, 5, 4, 5],
    [6, 5, 3, 3, 3, 3, 3, 3, 4, 3, 6],
    [6, 5, 3, 3, 3, 2, 3, 4, 3, 6],
    [6, 4, 5, 3, 2, 2, 3, 3, 6],
    [6, 6, 4, 2, 2, 3, 2, 5],
    [5, 5, 3, 2, 2, 2, 4],
    [4, 4, 3, 3, 1, 3],
    [4, 4, 2, 1, 3],
    [3, 3, 1, 2],
    [2, 2, 1],
    [1, 1]]
kTzCode = [
    [1, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 1],
    [7, 6, 5, 4, 3, 5, 4, 3, 2, 3, 2, 3, 2, 1, 0],
    [5, 7, 6, 5, 4, 3, 4, 3, 2, 3, 2, 1, 1, 0],
    [3, 7, 5, 4, 6, 5, 4, 3, 3, 2, 2, 1, 0],
    [5, 4, 3, 7, 6, 5, 4, 3, 2, 1, 1, 0],
    [1, 1, 7, 6, 5, 4, 3, 2, 1, 1, 0],
    [1, 1, 5, 4, 3, 3, 2, 1, 1, 0],
    [1, 1, 1, 3, 3, 2, 2, 1, 0],
    [1, 0, 1, 3, 2, 1, 1, 1],
    [1, 0, 1, 3, 2, 1, 1],
    [0, 1, 1, 2, 1, 3],
    [0, 1, 1, 1, 1],
    [0, 1, 1, 1],
    [0, 1, 1],
    [0, 1]]
# Table 9-9 total_zeros for chroma DC 2x2
kTzDcLen = [[1, 2, 3, 3], [1, 2, 2], [1, 1]]
kTzDcCode = [[1, 1, 1, 0], [1, 1, 0], [1, 0]]
# Table 9-10 run_before [min(zerosLeft,7)-1][run_before]
kRunLen = [[1, 1], [1, 2, 2], [2, 2, 2, 2], [2, 2, 2, 3, 3], [2, 2, 3, 3, 3, 3], [2, 3, 3, 3, 3, 3, 3],
           [3, 3, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9, 10, 11]]
kRunCode = [[1, 0], [1, 1, 0], [3, 2, 1, 0], [3, 2, 1, 1, 0], [3, 2, 3, 2, 1, 0], [3, 0, 1, 3, 2, 5, 4],
            [7, 6, 5, 4, 3, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1]]
# Table 9-4: coded_block_pattern -> codeNum, {intra, inter}
_kCbp = [(47, 0), (31, 16), (15, 1), (0, 2), (23, 4), (27, 8), (29, 32), (30, 3), (7, 5), (11, 10),
         (13, 12), (14, 15), (39, 47), (43, 7), (45, 11), (46, 13), (16, 14), (3, 6), (5, 9), (10, 31),
         (12, 35), (19, 37), (21, 42), (26, 44), (28, 33), (35, 34), (37, 36), (42, 40), (44, 39), (1, 43),
         (2, 45), (4, 46), (8, 17), (17, 18), (18, 20), (20, 24), (24, 19), (6, 21), (9, 26), (22, 28),
         (25, 23), (32, 27), (33, 29), (34, 30), (36, 22), (40, 25), (38, 38), (41, 41)]
CBP_CODE_INTRA = {c[0]: i for i, c in enumerate(_kCbp)}
CBP_CODE_INTER = {c[1]: i for i, c in enumerate(_kCbp)}
# Table 8-15
kQpC = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
        29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39]
# 4x4 luma block index (clause 6.4.3) <-> position inside the macroblock
BLK_X = [0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3]
BLK_Y = [0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3]
BLK_AT = [[0, 1, 4, 5], [2, 3, 6, 7], [8, 9, 12, 13], [10, 11, 14, 15]]   # [y][x]
# Table A-1: level_idc -> (MaxDPB bytes, MaxFS) as the reference uses them (h264bsd_seq_param_set.c:380-488)
LEVELS = {10: (152064, 99), 11: (345600, 396), 12: (912384, 396), 13: (912384, 396), 20: (912384, 396),
          21: (1824768, 792), 22: (3110400, 1620), 30: (3110400, 1620), 31: (6912000, 3600), 40: (12582912, 8192)}

INTER, I4, I16, PCM = 0, 1, 2, 3


# ---------------------------------------------------------------------------------------------- bits
class BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def u(self, n, v):
        assert 0 <= v < (1 << n), (n, v)
        self.acc = (self.acc << n) | v
        self.n += n
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def ue(self, v):
        assert v >= 0
        v += 1
        n = v.bit_length()
        self.u(2 * n - 1, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def te(self, v, rng):
        if rng > 1:
            self.ue(v)
        else:
            self.u(1, 1 - v)

    def aligned(self):
        return self.n == 0

    def trailing(self):   # rbsp_trailing_bits
        self.u(1, 1)
        while self.n:
            self.u(1, 0)

    def bitpos(self):
        return len(self.out) * 8 + self.n


def nal(ref_idc, typ, rbsp):
    """Annex B: start code + NAL header + RBSP with emulation prevention (clause 7.4.1.1)"""
    out = bytearray(b"\x00\x00\x00\x01")
    out.append((ref_idc << 5) | typ)
    zeros = 0
    for b in rbsp:
        if zeros >= 2 and b <= 3:
            out.append(3)
            zeros = 0
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


# ---------------------------------------------------------------------------------------------- CAVLC (clause 9.2)
def cavlc_block(bw, coef, nC):
    """residual_block_cavlc for `coef` (scan order, len 4 / 15 / 16); returns TotalCoeff"""
    maxn = len(coef)
    nz = [i for i, c in enumerate(coef) if c]
    tc = len(nz)
    t1 = 0
    for i in reversed(nz):
        if abs(coef[i]) == 1 and t1 < 3:
            t1 += 1
        else:
            break
    if nC < 0:
        bw.u(kTokDcLen[t1][tc], kTokDcCode[t1][tc])
    elif nC >= 8:
        bw.u(6, 3 if tc == 0 else ((tc - 1) << 2) | t1)
    else:
        t = 0 if nC < 2 else 1 if nC < 4 else 2
        bw.u(kTokLen[t][t1][tc], kTokCode[t][t1][tc])
    if tc == 0:
        return 0
    lv = [coef[i] for i in reversed(nz)]          # highest frequency first
    for v in lv[:t1]:
        bw.u(1, 1 if v < 0 else 0)
    sl = 1 if (tc > 10 and t1 < 3) else 0
    for k in range(t1, tc):
        v = lv[k]
        code = 2 * abs(v) - 2 if v > 0 else 2 * abs(v) - 1
        if k == t1 and t1 < 3:
            code -= 2
        if sl == 0:
            if code < 14:
                bw.u(code + 1, 1)
            elif code < 30:
                bw.u(15, 1)
                bw.u(4, code - 14)
            else:
                assert code - 30 < 4096
                bw.u(16, 1)
                bw.u(12, code - 30)
        else:
            pre = code >> sl
            if pre < 15:
                bw.u(pre + 1, 1)
                bw.u(sl, code & ((1 << sl) - 1))
            else:
                assert code - (15 << sl) < 4096
                bw.u(16, 1)
                bw.u(12, code - (15 << sl))
        if sl == 0:
            sl = 1
        if abs(v) > (3 << (sl - 1)) and sl < 6:
            sl += 1
    total_zeros = nz[-1] + 1 - tc
    if tc < maxn:
        if maxn == 4:
            bw.u(kTzDcLen[tc - 1][total_zeros], kTzDcCode[tc - 1][total_zeros])
        else:
            bw.u(kTzLen[tc - 1][total_zeros], kTzCode[tc - 1][total_zeros])
    zl = total_zeros
    for k in range(tc - 1):
        if zl <= 0:
            break
        run = nz[tc - 1 - k] - nz[tc - 2 - k] - 1
        row = min(zl, 7) - 1
        bw.u(kRunLen[row][run], kRunCode[row][run])
        zl -= run
    return tc


# ---------------------------------------------------------------------------------------------- slice group maps (8.2.2)
def slice_group_map(fmo, W, H, change_cycle):
    size = W * H
    n = fmo["groups"]
    if n == 1:
        return [0] * size
    t = fmo["type"]
    m = [0] * size
    if t == 0:
        i = 0
        while i < size:
            for g in range(n):
                for j in range(fmo["run"][g]):
                    if i + j < size:
                        m[i + j] = g
                i += fmo["run"][g]
                if i >= size:
                    break
    elif t == 1:
        for i in range(size):
            m[i] = ((i % W) + (((i // W) * n) // 2)) % n
    elif t == 2:
        m = [n - 1] * size
        for g in range(n - 2, -1, -1):
            tl, br = fmo["rect"][g]
            for y in range(tl // W, br // W + 1):
                for x in range(tl % W, br % W + 1):
                    m[y * W + x] = g
    elif t in (3, 4, 5):
        d = fmo["dir"]
        units0 = min(change_cycle * fmo["rate"], size)
        upper = size - units0 if d else units0
        if t == 3:
            m = [1] * size
            x = (W - d) // 2
            y = (H - d) // 2
            lb, tb, rb, bb = x, y, x, y
            xd, yd = d - 1, d
            k = 0
            while k < units0:
                vac = m[y * W + x] == 1
                if vac:
                    m[y * W + x] = 0
                if xd == -1 and x == lb:
                    lb = max(lb - 1, 0); x = lb; xd = 0; yd = 2 * d - 1
                elif xd == 1 and x == rb:
                    rb = min(rb + 1, W - 1); x = rb; xd = 0; yd = 1 - 2 * d
                elif yd == -1 and y == tb:
                    tb = max(tb - 1, 0); y = tb; xd = 1 - 2 * d; yd = 0
                elif yd == 1 and y == bb:
                    bb = min(bb + 1, H - 1); y = bb; xd = 2 * d - 1; yd = 0
                else:
                    x += xd; y += yd
                k += 1 if vac else 0
        elif t == 4:
            for i in range(size):
                m[i] = d if i < upper else 1 - d
        else:
            k = 0
            for x in range(W):
                for y in range(H):
                    m[y * W + x] = d if k < upper else 1 - d
                    k += 1
    else:
        m = list(fmo["ids"])
    return m


# ---------------------------------------------------------------------------------------------- picture model
class MbState:
    __slots__ = ("slice", "kind", "tc", "modes", "mv", "ref", "done")

    def __init__(self):
        self.slice = -1
        self.kind = INTER
        self.tc = [0] * 24
        self.modes = None
        self.mv = [[(0, 0)] * 4 for _ in range(4)]   # [y][x]
        self.ref = [0, 0, 0, 0]                      # per 8x8 quadrant (raster)
        self.done = None


def median(a, b, c):
    return a + b + c - max(a, b, c) - min(a, b, c)


class PictureWriter:
    """writes the slices of one picture; keeps the syntax-level state of its macroblocks"""

    def __init__(self, rng, W, H, pps, knobs):
        self.r = rng
        self.W, self.H = W, H
        self.pps = pps
        self.k = knobs
        self.mb = [MbState() for _ in range(W * H)]
        self.slice_counter = 0

    # -- neighbours
    def _nb(self, a, which, sid):
        x, y = a % self.W, a // self.W
        if which == 'A':
            n = a - 1 if x > 0 else -1
        elif which == 'B':
            n = a - self.W if y > 0 else -1
        elif which == 'C':
            n = a - self.W + 1 if (y > 0 and x < self.W - 1) else -1
        else:
            n = a - self.W - 1 if (y > 0 and x > 0) else -1
        if n >= 0 and self.mb[n].slice == sid:
            return n
        return -1

    def _nc(self, a, blk, tc_cur, sid):
        """nC of clause 9.2.1 for block `blk` (0..15 luma, 16..19 Cb, 20..23 Cr)"""
        A, B = self._nb(a, 'A', sid), self._nb(a, 'B', sid)
        if blk < 16:
            x, y = BLK_X[blk], BLK_Y[blk]
            na = tc_cur[BLK_AT[y][x - 1]] if x > 0 else (self.mb[A].tc[BLK_AT[y][3]] if A >= 0 else None)
            nb = tc_cur[BLK_AT[y - 1][x]] if y > 0 else (self.mb[B].tc[BLK_AT[3][x]] if B >= 0 else None)
        else:
            base = 16 if blk < 20 else 20
            i = blk - base
            cx, cy = i & 1, i >> 1
            na = tc_cur[base + cy * 2] if cx > 0 else (self.mb[A].tc[base + cy * 2 + 1] if A >= 0 else None)
            nb = tc_cur[base + cx] if cy > 0 else (self.mb[B].tc[base + 2 + cx] if B >= 0 else None)
        if na is not None and nb is not None:
            return (na + nb + 1) >> 1
        if na is not None:
            return na
        if nb is not None:
            return nb
        return 0

    # -- motion vector prediction (8.4.1.3)
    def _mv_nb(self, a, x, y, sid):
        """(available, refIdx, mv) of the 4x4 block at (x, y) relative to macroblock a"""
        if 0 <= x < 4 and 0 <= y < 4:
            m = self.mb[a]
            if not m.done[y][x]:
                return (False, -1, (0, 0))
            return (True, m.ref[(y >> 1) * 2 + (x >> 1)], m.mv[y][x])
        if y >= 0 and x > 3:
            return (False, -1, (0, 0))
        if y < 0:
            n = self._nb(a, 'D' if x < 0 else 'C' if x > 3 else 'B', sid)
        else:
            n = self._nb(a, 'A', sid)
        if n < 0:
            return (False, -1, (0, 0))
        m = self.mb[n]
        if m.kind != INTER:
            return (True, -1, (0, 0))
        xx, yy = x & 3, y & 3
        return (True, m.ref[(yy >> 1) * 2 + (xx >> 1)], m.mv[yy][xx])

    def _pred_mv(self, a, x, y, w, ref, hint, sid):
        A = self._mv_nb(a, x - 1, y, sid)
        B = self._mv_nb(a, x, y - 1, sid)
        C = self._mv_nb(a, x + w, y - 1, sid)
        if not C[0]:
            C = self._mv_nb(a, x - 1, y - 1, sid)
        if not B[0] and not C[0] and A[0]:
            B = C = A
        if hint == 'B' and B[1] == ref:
            return B[2]
        if hint == 'A' and A[1] == ref:
            return A[2]
        if hint == 'C' and C[1] == ref:
            return C[2]
        match = [n for n in (A, B, C) if n[1] == ref]
        if len(match) == 1:
            return match[0][2]
        return (median(A[2][0], B[2][0], C[2][0]), median(A[2][1], B[2][1], C[2][1]))

    def _set_mv(self, m, x, y, w, h, mv, ref):
        for yy in range(y, y + h):
            for xx in range(x, x + w):
                m.mv[yy][xx] = mv
                m.done[yy][xx] = True

    def _pick_mv(self, mbx, mby):
        r = self.r
        if self.k.get("still") and r.random() < 0.92:      # a still scene: long runs of zero-vector copies
            return (0, 0)
        c = r.random()
        if c < 0.25:
            return (0, 0)
        if c < 0.40:
            return (4 * r.randint(-3, 3), 4 * r.randint(-3, 3))           # integer sample
        if c < 0.75:
            return (r.randint(-24, 24), r.randint(-24, 24))
        if c < 0.90:                                                       # around / across the picture edges
            px, py = 64 * self.W, 64 * self.H
            return (max(-8192, min(8191, r.choice([-1, 1]) * r.randint(0, px + 90) - (64 * mbx if r.random() < .5 else 0))),
                    max(-2048, min(2047, r.choice([-1, 1]) * r.randint(0, py + 90))))
        if c < 0.97:
            return (r.randint(-8192, 8191), r.randint(-2048, 2047))
        return (r.choice([-8192, 8191, -8191, 0]), r.choice([-2048, 2047, 0]))

    # -- residual
    def _block(self, n, budget, dense):
        """n coefficients in scan order, sum of magnitudes <= budget"""
        r = self.r
        coef = [0] * n
        if budget < 1:
            return coef
        c = r.random()
        cnt = 0 if c < 0.15 else r.randint(1, 3) if c < 0.6 else r.randint(1, n) if not dense else n
        cnt = min(cnt, n, int(budget))
        if cnt == 0:
            return coef
        # low frequencies more likely, sometimes anywhere
        if r.random() < 0.6:
            pos = sorted(r.sample(range(min(n, cnt + 3)), cnt))
        else:
            pos = sorted(r.sample(range(n), cnt))
        left = int(budget) - cnt
        for p in pos:
            c = r.random()
            extra = 0
            if left > 0:
                if c < 0.5:
                    extra = 0
                elif c < 0.85:
                    extra = r.randint(0, min(left, 3))
                elif c < 0.97:
                    extra = r.randint(0, min(left, 40))
                else:
                    extra = r.randint(0, min(left, 2060))
            left -= extra
            coef[p] = (1 + extra) * r.choice([-1, 1])
        return coef

    def _budgets(self, qp):
        qpc = kQpC[max(0, min(51, qp + self.pps["chroma_qp_offset"]))]
        return {"ac": 16000 // (29 << (qp // 6)), "dc": (16000 * 4) // (18 << (qp // 6)),
                "cac": 16000 // (29 << (qpc // 6)), "cdc": (16000 * 2) // (18 << (qpc // 6))}

    # -- one slice
    def write_slice(self, bw, mbs, is_p, qp, num_ref_active, valid_refs):
        """slice_data() for the macroblocks `mbs` (decoding order); updates the model.  valid_refs: the ref_idx values that
        point at a real picture (None: none does -> no inter prediction in this slice)"""
        r = self.r
        self.slice_counter += 1
        sid = self.slice_counter
        skip_run = 0
        wrote_any = False
        i = 0
        n = len(mbs)
        p_skip_prob = r.choice([0.0, 0.2, 0.5, 0.8]) if (is_p and valid_refs and 0 in valid_refs) else 0
        if p_skip_prob and self.k.get("still"):
            p_skip_prob = 0.85
        while i < n:
            a = mbs[i]
            m = self.mb[a]
            m.slice = sid
            if is_p and r.random() < p_skip_prob:
                self._skip_mb(a, sid)
                skip_run += 1
                i += 1
                continue
            if is_p:
                bw.ue(skip_run)
                skip_run = 0
            qp = self._coded_mb(bw, a, sid, is_p, qp, num_ref_active, valid_refs)
            wrote_any = True
            i += 1
        if is_p and skip_run:
            bw.ue(skip_run)
        bw.trailing()

    def _skip_mb(self, a, sid):
        m = self.mb[a]
        m.kind = INTER
        m.tc = [0] * 24
        m.modes = None
        m.ref = [0, 0, 0, 0]
        m.done = [[False] * 4 for _ in range(4)]
        A = self._mv_nb(a, -1, 0, sid)
        B = self._mv_nb(a, 0, -1, sid)
        if (not A[0]) or (not B[0]) or (A[1] == 0 and A[2] == (0, 0)) or (B[1] == 0 and B[2] == (0, 0)):
            mv = (0, 0)
        else:
            mv = self._pred_mv(a, 0, 0, 4, 0, None, sid)
        self._set_mv(m, 0, 0, 4, 4, mv, 0)

    def _intra_avail(self, a, sid):
        """availability of macroblocks A, B, C, D for intra prediction (incl. constrained_intra_pred)"""
        out = {}
        for w in "ABCD":
            n = self._nb(a, w, sid)
            if n >= 0 and self.pps["constrained_intra"] and self.mb[n].kind == INTER:
                n = -1
            out[w] = n
        return out

    def _coded_mb(self, bw, a, sid, is_p, qp, num_ref_active, valid_refs):
        r = self.r
        m = self.mb[a]
        mbx, mby = a % self.W, a // self.W
        off = 5 if is_p else 0
        if is_p and valid_refs and r.random() < self.k["p_inter"]:
            kind = INTER
        else:
            c = r.random()
            kind = PCM if c < self.k["pcm"] else I16 if c < 0.5 else I4
        tc = [0] * 24
        m.modes = None
        m.done = [[False] * 4 for _ in range(4)]
        i16_dc = None
        if kind == PCM:
            bw.ue(off + 25)
            while not bw.aligned():
                bw.u(1, 0)
            flat = r.random() < 0.3
            v = r.randint(0, 255)
            for _ in range(384):
                bw.u(8, v if flat else r.randint(0, 255))
            m.kind = PCM
            m.tc = [16] * 24
            return qp     # QP'Y of an I_PCM macroblock is 0 for the filter, the running QP is unchanged
        if kind == INTER:
            shape = r.choice([0, 0, 1, 2, 3, 3, 4]) if num_ref_active >= 1 else 0
            if shape == 4 and (r.random() < 0.5 or 0 not in valid_refs):
                shape = 3
            bw.ue(shape)
            m.kind = INTER

            def pick_ref():
                return valid_refs[0] if r.random() < 0.5 else r.choice(valid_refs)

            def put_ref(v):
                if num_ref_active > 1:
                    bw.te(v, num_ref_active - 1)

            if shape <= 2:
                parts = [(0, 0, 4, 4, None)] if shape == 0 else \
                        [(0, 0, 4, 2, 'B'), (0, 2, 4, 2, 'A')] if shape == 1 else [(0, 0, 2, 4, 'A'), (2, 0, 2, 4, 'C')]
                refs = [pick_ref() for _ in parts]
                for v in refs:
                    put_ref(v)
                for (x, y, w, h, hint), ref in zip(parts, refs):
                    for q in range(4):
                        qx, qy = (q & 1) * 2, (q >> 1) * 2
                        if x <= qx < x + w and y <= qy < y + h:
                            m.ref[q] = ref
                    pred = self._pred_mv(a, x, y, w, ref, hint, sid)
                    mv = self._pick_mv(mbx, mby)
                    bw.se(mv[0] - pred[0])
                    bw.se(mv[1] - pred[1])
                    self._set_mv(m, x, y, w, h, mv, ref)
            else:
                subs = [r.randrange(4) for _ in range(4)]
                for s in subs:
                    bw.ue(s)
                refs = [0 if shape == 4 else pick_ref() for _ in range(4)]
                if shape != 4:
                    for v in refs:
                        put_ref(v)
                m.ref = list(refs)
                for q in range(4):
                    qx, qy = (q & 1) * 2, (q >> 1) * 2
                    s = subs[q]
                    sub_parts = [(qx, qy, 2, 2)] if s == 0 else [(qx, qy, 2, 1), (qx, qy + 1, 2, 1)] if s == 1 else \
                                [(qx, qy, 1, 2), (qx + 1, qy, 1, 2)] if s == 2 else \
                                [(qx, qy, 1, 1), (qx + 1, qy, 1, 1), (qx, qy + 1, 1, 1), (qx + 1, qy + 1, 1, 1)]
                    for (x, y, w, h) in sub_parts:
                        pred = self._pred_mv(a, x, y, w, refs[q], None, sid)
                        mv = self._pick_mv(mbx, mby)
                        bw.se(mv[0] - pred[0])
                        bw.se(mv[1] - pred[1])
                        self._set_mv(m, x, y, w, h, mv, refs[q])
            cbp_l = r.choice([0, 0, 15, r.randrange(16)])
            cbp_c = r.choice([0, 0, 1, 2])
            if self.k.get("still") and r.random() < 0.7:
                cbp_l = cbp_c = 0
            cbp = cbp_l | (cbp_c << 4)
            bw.ue(CBP_CODE_INTER[cbp])
        else:
            av = self._intra_avail(a, sid)
            hasA, hasB, hasD = av['A'] >= 0, av['B'] >= 0, av['D'] >= 0
            m.kind = kind
            m.ref = [0, 0, 0, 0]
            m.mv = [[(0, 0)] * 4 for _ in range(4)]
            chroma_ok = [0] + ([1] if hasA else []) + ([2] if hasB else []) + ([3] if hasA and hasB and hasD else [])
            chroma_mode = r.choice(chroma_ok)
            if kind == I4:
                bw.ue(off + 0)
                modes = [2] * 16
                for blk in range(16):
                    x, y = BLK_X[blk], BLK_Y[blk]
                    # neighbouring blocks A / B: availability (for intra prediction) and their mode
                    if x > 0:
                        bA, mA = True, modes[BLK_AT[y][x - 1]]
                    else:
                        bA = hasA
                        mA = self.mb[av['A']].modes[BLK_AT[y][3]] if (bA and self.mb[av['A']].kind == I4) else 2
                    if y > 0:
                        bB, mB = True, modes[BLK_AT[y - 1][x]]
                    else:
                        bB = hasB
                        mB = self.mb[av['B']].modes[BLK_AT[3][x]] if (bB and self.mb[av['B']].kind == I4) else 2
                    bD = hasD if (x == 0 and y == 0) else hasA if x == 0 else hasB if y == 0 else True
                    pred = 2 if not (bA and bB) else min(mA, mB)
                    ok = [2]
                    if bB:
                        ok += [0, 3, 7]
                    if bA:
                        ok += [1, 8]
                    if bA and bB and bD:
                        ok += [4, 5, 6]
                    mode = pred if (pred in ok and r.random() < 0.3) else r.choice(ok)
                    if mode == pred:
                        bw.u(1, 1)
                    else:
                        bw.u(1, 0)
                        bw.u(3, mode if mode < pred else mode - 1)
                    modes[blk] = mode
                m.modes = modes
                bw.ue(chroma_mode)
                cbp_l = r.choice([0, 15, r.randrange(16), r.randrange(16)])
                cbp_c = r.choice([0, 1, 2, 2])
                cbp = cbp_l | (cbp_c << 4)
                bw.ue(CBP_CODE_INTRA[cbp])
            else:
                ok = [2] + ([0] if hasB else []) + ([1] if hasA else []) + ([3] if hasA and hasB and hasD else [])
                pm = r.choice(ok)
                cbp_c = r.choice([0, 1, 2])
                ac = r.random() < 0.5
                cbp_l = 15 if ac else 0
                cbp = cbp_l | (cbp_c << 4)
                bw.ue(off + 1 + pm + 4 * cbp_c + (12 if ac else 0))
                bw.ue(chroma_mode)
        if cbp or kind == I16:
            # mb_qp_delta: mostly small, sometimes to the ends of the range / wrapping around (7.4.5)
            c = r.random()
            if c < 0.5:
                d = 0
            elif c < 0.9:
                d = r.randint(-3, 3)
            else:
                d = r.randint(-26, 25)
            if self.k["qp_lo"] is not None:      # keep the quantiser inside a band (large coefficients need a small one)
                tgt = max(self.k["qp_lo"], min(self.k["qp_hi"], qp + d))
                d = tgt - qp
            bw.se(d)
            qp = (qp + d + 52) % 52
            bud = self._budgets(qp)
            dense = r.random() < self.k["dense"]
            if kind == I16:
                dcc = self._block(16, bud["dc"], dense)
                cavlc_block(bw, dcc, self._nc(a, 0, tc, sid))
            for blk in range(16):
                if cbp_l & (1 << (blk >> 2)):
                    nC = self._nc(a, blk, tc, sid)
                    if kind == I16:
                        tc[blk] = cavlc_block(bw, self._block(15, bud["ac"], dense), nC)
                    else:
                        tc[blk] = cavlc_block(bw, self._block(16, bud["ac"] * 2 if kind != I16 else bud["ac"], dense), nC)
            if cbp_c:
                for _ in range(2):
                    cavlc_block(bw, self._block(4, bud["cdc"], dense), -1)
            if cbp_c == 2:
                for blk in range(16, 24):
                    tc[blk] = cavlc_block(bw, self._block(15, bud["cac"], dense), self._nc(a, blk, tc, sid))
        m.tc = tc
        return qp


# ---------------------------------------------------------------------------------------------- parameter sets
def write_vui(bw, r, sps):
    if r.random() < 0.7:
        bw.u(1, 1)
        idc = r.choice([1, 2, 13, 255])
        bw.u(8, idc)
        if idc == 255:
            bw.u(16, r.randint(1, 65535)); bw.u(16, r.randint(1, 65535))
    else:
        bw.u(1, 0)
    if r.random() < 0.5:
        bw.u(1, 1); bw.u(1, r.randint(0, 1))          # overscan
    else:
        bw.u(1, 0)
    if r.random() < 0.7:                                # video_signal_type
        bw.u(1, 1)
        bw.u(3, r.randint(0, 5)); bw.u(1, r.randint(0, 1))
        if r.random() < 0.6:
            bw.u(1, 1); bw.u(8, r.choice([1, 2, 5, 6])); bw.u(8, r.choice([1, 2, 6])); bw.u(8, r.choice([1, 2, 5, 6]))
        else:
            bw.u(1, 0)
    else:
        bw.u(1, 0)
    if r.random() < 0.4:
        bw.u(1, 1); bw.ue(r.randint(0, 5)); bw.ue(r.randint(0, 5))   # chroma_loc
    else:
        bw.u(1, 0)
    if r.random() < 0.5:
        bw.u(1, 1); bw.u(32, r.randint(1, 100000)); bw.u(32, r.randint(1, 100000)); bw.u(1, r.randint(0, 1))
    else:
        bw.u(1, 0)
    hrd = 0
    for _ in range(2):                                  # nal / vcl hrd_parameters
        if r.random() < 0.3:
            hrd = 1
            bw.u(1, 1)
            cnt = r.randint(1, 3)
            bw.ue(cnt - 1); bw.u(4, r.randint(0, 15)); bw.u(4, r.randint(0, 15))
            for _i in range(cnt):
                bw.ue(r.randint(0, 100000)); bw.ue(r.randint(0, 100000)); bw.u(1, r.randint(0, 1))
            for _i in range(4):
                bw.u(5, r.randint(0, 31))
        else:
            bw.u(1, 0)
    if hrd:
        bw.u(1, r.randint(0, 1))                        # low_delay_hrd_flag
    bw.u(1, r.randint(0, 1))                            # pic_struct_present_flag
    if r.random() < 0.6:                                # bitstream_restriction
        bw.u(1, 1)
        bw.u(1, r.randint(0, 1)); bw.ue(r.randint(0, 16)); bw.ue(r.randint(0, 16)); bw.ue(r.randint(0, 16)); bw.ue(r.randint(0, 16))
        mdfb = r.randint(max(sps["num_ref_frames"], 0), sps["dpb_size"])
        nrf = r.randint(0, mdfb)
        bw.ue(nrf); bw.ue(mdfb)
        sps["dpb_size"] = max(1, mdfb)
    else:
        bw.u(1, 0)


def write_sps(r, sps):
    bw = BitWriter()
    bw.u(8, 66)
    bw.u(1, 1); bw.u(1, r.randint(0, 1)); bw.u(1, 0); bw.u(5, 0)
    bw.u(8, sps["level"])
    bw.ue(sps["id"])
    bw.ue(sps["log2_max_frame_num"] - 4)
    bw.ue(sps["poc_type"])
    if sps["poc_type"] == 0:
        bw.ue(sps["log2_max_poc_lsb"] - 4)
    elif sps["poc_type"] == 1:
        bw.u(1, sps["delta_always_zero"])
        bw.se(sps["offset_non_ref"])
        bw.se(sps["offset_top_bottom"])
        bw.ue(len(sps["offset_ref_frame"]))
        for v in sps["offset_ref_frame"]:
            bw.se(v)
    bw.ue(sps["num_ref_frames"])
    bw.u(1, sps["gaps_allowed"])
    bw.ue(sps["W"] - 1)
    bw.ue(sps["H"] - 1)
    bw.u(1, 1)                      # frame_mbs_only_flag
    bw.u(1, r.randint(0, 1))        # direct_8x8_inference_flag
    if sps["crop"]:
        bw.u(1, 1)
        for v in sps["crop"]:
            bw.ue(v)
    else:
        bw.u(1, 0)
    if sps["vui"]:
        bw.u(1, 1)
        write_vui(bw, r, sps)
    else:
        bw.u(1, 0)
    bw.trailing()
    return nal(r.choice([1, 3]), 7, bw.out)


def write_pps(r, pps, size):
    bw = BitWriter()
    bw.ue(pps["id"])
    bw.ue(pps["sps_id"])
    bw.u(1, 0)                                  # entropy_coding_mode_flag
    bw.u(1, pps["pic_order_present"])
    fmo = pps["fmo"]
    bw.ue(fmo["groups"] - 1)
    if fmo["groups"] > 1:
        bw.ue(fmo["type"])
        if fmo["type"] == 0:
            for g in range(fmo["groups"]):
                bw.ue(fmo["run"][g] - 1)
        elif fmo["type"] == 2:
            for g in range(fmo["groups"] - 1):
                bw.ue(fmo["rect"][g][0]); bw.ue(fmo["rect"][g][1])
        elif fmo["type"] in (3, 4, 5):
            bw.u(1, fmo["dir"])
            bw.ue(fmo["rate"] - 1)
        elif fmo["type"] == 6:
            bw.ue(size - 1)
            nb = (fmo["groups"] - 1).bit_length()
            for v in fmo["ids"]:
                bw.u(nb, v)
    bw.ue(pps["num_ref_idx_default"] - 1)
    bw.ue(r.randint(0, 3))                      # num_ref_idx_l1_default_active_minus1 (unused in Baseline)
    bw.u(1, 0); bw.u(2, 0)                      # weighted_pred_flag, weighted_bipred_idc
    bw.se(pps["pic_init_qp"] - 26)
    bw.se(r.randint(-26, 25))                   # pic_init_qs_minus26
    bw.se(pps["chroma_qp_offset"])
    bw.u(1, pps["deblock_ctrl"])
    bw.u(1, pps["constrained_intra"])
    bw.u(1, pps["redundant_present"])
    bw.trailing()
    return nal(r.choice([1, 2, 3]), 8, bw.out)


def random_fmo(r, W, H, allow):
    size = W * H
    if not allow or size < 2 or r.random() < 0.5:
        return {"groups": 1}
    t = r.randrange(7)
    if t in (3, 4, 5):
        return {"groups": 2, "type": t, "dir": r.randint(0, 1), "rate": r.randint(1, size)}
    n = r.randint(2, min(8, size))
    fmo = {"groups": n, "type": t}
    if t == 0:
        fmo["run"] = [r.randint(1, max(1, size // 2)) for _ in range(n)]
    elif t == 2:
        rect = []
        for _ in range(n - 1):
            x0, y0 = r.randrange(W), r.randrange(H)
            x1, y1 = r.randint(x0, W - 1), r.randint(y0, H - 1)
            rect.append((y0 * W + x0, y1 * W + x1))
        fmo["rect"] = rect
    elif t == 6:
        fmo["ids"] = [r.randrange(n) for _ in range(size)]
    return fmo


# ---------------------------------------------------------------------------------------------- stream
def _picnum(f, frame_num, max_fn):
    return f["fn"] if f["fn"] <= frame_num else f["fn"] - max_fn


def _sliding_window(refs, nrf, frame_num, max_fn):
    """8.2.5.3: with all reference frames in use the short-term frame with the smallest FrameNumWrap goes"""
    if len(refs) >= max(nrf, 1):
        st = [f for f in refs if f["lt"] is None]
        if st:
            refs.remove(min(st, key=lambda f: _picnum(f, frame_num, max_fn)))


def _random_mmco(r, refs, max_lt, nrf, frame_num, max_fn, must_drop=()):
    """memory_management_control_operation list (7.3.3.3 / 8.2.5.4) that leaves room for the current picture and
    un-marks the short-term frames `must_drop`.  Returns (ops, refs after them, max long-term index after them,
    long-term index of the current picture or None, had operation 5)."""
    keep = [f for f in refs if not any(f is d for d in must_drop)]
    ops = [(1, frame_num - _picnum(f, frame_num, max_fn) - 1) for f in must_drop]
    refs = [dict(f) for f in keep]
    cur_lt = None
    had5 = False
    while len(refs) >= max(nrf, 1):          # room for the current picture (none of the operations below adds a frame)
        f = r.choice(refs)
        ops.append((1, frame_num - _picnum(f, frame_num, max_fn) - 1) if f["lt"] is None else (2, f["lt"]))
        refs.remove(f)

    def drop_lt(idx):
        for f in [f for f in refs if f["lt"] == idx]:
            refs.remove(f)

    for _ in range(r.randint(1, 4)):
        st = [f for f in refs if f["lt"] is None]
        lt = [f for f in refs if f["lt"] is not None]
        # the reference accepts at most one each of operations 4, 5, 6 and not 5 together with 1..3
        # (h264bsd_slice_header.c:697-699)
        kinds = [o[0] for o in ops]
        choices = [] if 4 in kinds else [4]
        if not had5:
            st_real = [f for f in st if not f.get("ne")]      # (a gap filler cannot become a long-term frame)
            choices += [1] * bool(st) + [2] * bool(lt) + [3] * bool(st_real and max_lt is not None)
            if r.random() < 0.15 and not ops and frame_num != 1:   # operation 5 only as the first one: what it does to a picture
                choices.append(5)                        # that operation 6 has already marked is not worth guessing;
                # not with frame_num 1: the next picture has frame_num 1 again and the reference tells pictures apart by
                # frame_num / POC syntax / nal_ref_idc only, not by pic_parameter_set_id (h264bsd_storage.c:626-745)
        if max_lt is not None and cur_lt is None:
            choices.append(6)
        if not choices:
            break
        op = r.choice(choices)
        if op == 1:
            f = r.choice(st)
            ops.append((1, frame_num - _picnum(f, frame_num, max_fn) - 1))
            refs.remove(f)
        elif op == 2:
            f = r.choice(lt)
            ops.append((2, f["lt"]))
            refs.remove(f)
        elif op == 3:
            f = r.choice(st_real)
            idx = r.randint(0, max_lt)
            ops.append((3, frame_num - _picnum(f, frame_num, max_fn) - 1, idx))
            for g in [g for g in refs if g["lt"] == idx and g is not f]:
                refs.remove(g)
            f["lt"] = idx
        elif op == 4:
            v = r.randint(0, min(nrf, 3))
            ops.append((4, v))
            for g in [g for g in refs if g["lt"] is not None and g["lt"] >= v]:
                refs.remove(g)
            max_lt = v - 1 if v else None
        elif op == 5:
            ops.append((5,))
            refs.clear()
            max_lt = None
            had5 = True
            frame_num = 0
        else:
            idx = r.randint(0, max_lt)
            ops.append((6, idx))
            drop_lt(idx)
            cur_lt = idx
            break
    return ops, refs, max_lt, cur_lt, had5


def make_stream(seed, **force):
    """One random valid stream.  `force` overrides knobs: W, H, pictures, fmo (bool), multi_slice (bool), aso (bool),
    num_ref_frames, poc_type, i_only (bool), dense (float), vui (bool), mmco (bool), gaps (bool), redundant (bool),
    still (bool: mostly zero vectors and P_Skip -- long runs of plain copies), resend (bool: parameter sets repeated / changed
    between pictures)."""
    r = random.Random(seed)
    out = bytearray()

    # ---- sequence
    W = force.get("W") or r.choice([1, 2, 3, 4, 5, 6, 8, 11])
    H = force.get("H") or r.choice([1, 2, 3, 4, 5, 6, 9])
    size = W * H
    level = r.choice([l for l, (dpb, fs) in LEVELS.items() if fs >= size and l < 40] or [40])     # (level 4 only for pictures that need it)
    dpb_size = min(LEVELS[level][0] // (size * 384), 16)
    i_only = force.get("i_only", r.random() < 0.08)
    nrf = force.get("num_ref_frames")
    if nrf is None:
        nrf = 0 if i_only and r.random() < 0.5 else r.randint(1, min(dpb_size, 5))
    nrf = min(nrf, dpb_size)
    sps = {"id": r.choice([0, 0, 1, 31]), "level": level, "log2_max_frame_num": r.choice([4, 4, 5, 8, 16]),
           "poc_type": force.get("poc_type", r.choice([0, 0, 1, 2, 2])), "log2_max_poc_lsb": r.choice([4, 5, 8, 16]),
           "delta_always_zero": r.randint(0, 1), "offset_non_ref": r.randint(-3, 3), "offset_top_bottom": r.randint(-2, 2),
           "offset_ref_frame": [r.randint(1, 6) for _ in range(r.randint(0, 3))],
           "num_ref_frames": nrf, "gaps_allowed": 1 if (nrf >= 1 and force.get("gaps", r.random() < 0.2)) else 0,
           "W": W, "H": H, "dpb_size": dpb_size, "vui": force.get("vui", r.random() < 0.4), "crop": None}
    if r.random() < 0.4:
        cl, cr_ = r.randint(0, 3), r.randint(0, 3)
        ct, cb = r.randint(0, 3), r.randint(0, 3)
        if cl + cr_ < 8 * W and ct + cb < 8 * H:
            sps["crop"] = (cl, cr_, ct, cb)
    sps_nal = write_sps(r, sps)
    out += sps_nal
    max_fn = 1 << sps["log2_max_frame_num"]

    # ---- picture parameter sets
    allow_fmo = force.get("fmo", True)
    ppss = []
    for pid in r.sample(range(0, 256), r.randint(1, 3)):
        pps = {"id": pid, "sps_id": sps["id"], "pic_order_present": r.randint(0, 1), "fmo": random_fmo(r, W, H, allow_fmo),
               "num_ref_idx_default": r.randint(1, 4), "pic_init_qp": r.randint(10, 45), "chroma_qp_offset": r.randint(-12, 12),
               "deblock_ctrl": r.randint(0, 1), "constrained_intra": 1 if r.random() < 0.3 else 0,
               "redundant_present": 1 if force.get("redundant", r.random() < 0.25) else 0}
        ppss.append(pps)
        out += write_pps(r, pps, size)

    # ---- pictures
    n_pics = force.get("pictures") or r.randint(2, 9)
    if sps["log2_max_frame_num"] == 4 and r.random() < 0.3 and size <= 12:
        n_pics = r.randint(18, 40)                 # frame_num wraps around
    use_mmco = nrf >= 2 and force.get("mmco", r.random() < 0.4)
    refs = []           # reference frames in the DPB: {fn: frame_num, lt: long-term index or None, ne: "non-existing" (gap filler)}
    max_lt = None       # MaxLongTermFrameIdx ("no long-term frame indices" = None)
    prev_ref_fn = 0
    idr_id = r.randint(0, 100)
    poc_base = 0        # POC type 0: highest POC since the last IDR / operation 5
    poc_seen = []
    prev_was_nonref = False
    after5 = False      # the previous picture carried memory_management_control_operation 5
    knobs = {"p_inter": r.choice([0.5, 0.8, 0.95, 1.0]), "pcm": r.choice([0, 0.02, 0.1]),
             "dense": force.get("dense", r.choice([0.0, 0.1, 0.5])), "qp_lo": None, "qp_hi": None}
    if r.random() < 0.35:
        knobs["qp_lo"], knobs["qp_hi"] = r.choice([(0, 12), (20, 35), (40, 51), (0, 51)])
    knobs["still"] = force.get("still", False)
    resend = force.get("resend", r.random() < 0.3)
    for pic in range(n_pics):
        idr = pic == 0 or (r.random() < 0.1)
        is_ref = True if idr else (nrf > 0 and (r.random() < 0.8 or prev_was_nonref))
        # parameter sets again, as streams meant for random access carry them: the same SPS (no effect, h264bsdCompareSeqParamSets),
        # a picture parameter set with the same or with new content (takes effect with the next picture that names it)
        if resend and pic > 0 and r.random() < (0.6 if idr else 0.25):
            if r.random() < 0.5:
                out += sps_nal
            for q in ppss:
                if r.random() < 0.6:
                    if r.random() < 0.5:
                        q["pic_init_qp"] = r.randint(10, 45)
                        q["chroma_qp_offset"] = r.randint(-12, 12)
                        q["deblock_ctrl"] = r.randint(0, 1)
                        q["pic_order_present"] = r.randint(0, 1)
                        q["num_ref_idx_default"] = r.randint(1, 4)
                        q["constrained_intra"] = 1 if r.random() < 0.3 else 0
                        if allow_fmo and r.random() < 0.5:
                            q["fmo"] = random_fmo(r, W, H, True)
                    out += write_pps(r, q, size)
        pps = r.choice(ppss)
        gap = 0
        if idr:
            frame_num = 0
            idr_id = (idr_id + 1) % 65536
            poc_base = 0
            poc_seen = []
        else:
            st_now = [f for f in refs if f["lt"] is None]
            if sps["gaps_allowed"] and r.random() < 0.3 and (st_now or len(refs) < nrf) and not after5:
                gap = r.randint(1, 3)
            # the decoder fills a gap in frame_num with "non-existing" short-term frames, sliding window each (8.2.5.2)
            for g in range(gap):
                fn = (prev_ref_fn + 1 + g) % max_fn
                _sliding_window(refs, nrf, fn, max_fn)
                refs.append({"fn": fn, "lt": None, "ne": True})
            frame_num = (prev_ref_fn + 1 + gap) % max_fn
        can_p = (not idr) and (not i_only) and any(not f.get("ne") for f in refs)

        # initial reference list (8.2.4.2.1): short-term by PicNum descending, long-term by index ascending
        st = sorted([f for f in refs if f["lt"] is None], key=lambda f: _picnum(f, frame_num, max_fn), reverse=True)
        lt = sorted([f for f in refs if f["lt"] is not None], key=lambda f: f["lt"])
        init_list = st + lt

        # slices: every slice group's macroblocks (raster order) cut into runs
        fmo = pps["fmo"]
        cycle = 0
        if fmo["groups"] > 1 and fmo["type"] in (3, 4, 5):
            cycle = r.randint(0, (size + fmo["rate"] - 1) // fmo["rate"])
        sgm = slice_group_map(fmo, W, H, cycle)
        slices = []
        multi = force.get("multi_slice", r.random() < 0.6)
        for g in range(fmo["groups"]):
            mbs = [i for i in range(size) if sgm[i] == g]
            while mbs:
                k = len(mbs) if not multi else r.randint(1, len(mbs))
                slices.append(mbs[:k])
                mbs = mbs[k:]
        if force.get("aso", r.random() < 0.3):
            r.shuffle(slices)
        # POC type 0: mostly ascending, sometimes out of decoding order (output reordering)
        poc = 0
        if sps["poc_type"] == 0 and not idr:
            poc = poc_base + 2 * (r.choice([1, 2, 3]) if r.random() < 0.25 else 1)
            lower = [p for p in range(2, poc, 2) if p not in poc_seen]
            if lower and r.random() < 0.3:
                poc = r.choice(lower[-2:])
        poc_seen.append(poc)
        poc_base = max(poc_base, poc)
        poc1_delta = r.choice([0, 0, 0, 2, 4])
        # decoded reference picture marking, the same in every slice header of the picture
        lt_idr = idr and nrf >= 2 and r.random() < 0.3
        no_out = r.randint(0, 1)
        mmco = None
        # a short-term frame that sliding-window marking would long have dropped must go before frame_num comes round to
        # its own value again (frame numbers of the reference frames are distinct, 7.4.3)
        stale = [f for f in refs if f["lt"] is None and 0 < (f["fn"] - frame_num) % max_fn <= 4] if not idr else []
        if stale:
            is_ref = True
        # sliding-window marking needs a short-term frame to drop when all reference frames are in use (8.2.5.3)
        no_slide = len(refs) >= max(nrf, 1) and not any(f["lt"] is None for f in refs)
        if is_ref and not idr and ((use_mmco and r.random() < 0.5) or stale or no_slide):
            mmco = _random_mmco(r, refs, max_lt, nrf, frame_num, max_fn, stale)
        pw = PictureWriter(r, W, H, pps, knobs)
        if "trace" in force:
            force["trace"].append(dict(pic=pic, idr=idr, ref=is_ref, frame_num=frame_num, poc=poc, refs=[dict(f) for f in refs],
                                       mmco=mmco[0] if mmco else None, slices=len(slices), pps=pps["id"]))

        def emit_slice(mbs, redundant_cnt):
            is_p = can_p and r.random() < 0.8
            bw = BitWriter()
            bw.ue(mbs[0])
            bw.ue((0 if is_p else 2) + (5 if (r.random() < 0.2 and len(slices) == 1 and not pps["redundant_present"]) else 0))
            bw.ue(pps["id"])
            bw.u(sps["log2_max_frame_num"], frame_num)
            if idr:
                bw.ue(idr_id)
            if sps["poc_type"] == 0:
                bw.u(sps["log2_max_poc_lsb"], poc % (1 << sps["log2_max_poc_lsb"]))
                if pps["pic_order_present"]:
                    bw.se(0)
            elif sps["poc_type"] == 1 and not sps["delta_always_zero"]:
                # an IDR picture must come out with POC 0: min(d0, d0 + offset_for_top_to_bottom_field + d1) == 0
                # (the reference checks it, h264bsd_slice_header.c:234-243)
                bw.se(max(0, -sps["offset_top_bottom"]) if idr else poc1_delta)
                if pps["pic_order_present"]:
                    bw.se(0)
            if pps["redundant_present"]:
                bw.ue(redundant_cnt)
            num_active = pps["num_ref_idx_default"]
            cur_list = list(init_list)
            if is_p:
                if r.random() < 0.5:
                    num_active = r.randint(1, max(1, min(len(init_list), 8)))
                    bw.u(1, 1)
                    bw.ue(num_active - 1)
                else:
                    bw.u(1, 0)
                # ref_pic_list_reordering (7.3.3.1 / 8.2.4.3)
                real = [f for f in init_list if not f.get("ne")]       # (a gap filler cannot be named, h264bsd_dpb.c:288)
                if r.random() < 0.4 and num_active <= len(init_list) and real:
                    bw.u(1, 1)
                    pred = frame_num
                    for idx in range(r.randint(1, min(num_active, 3))):
                        f = r.choice(real)
                        if f["lt"] is None:
                            pn = _picnum(f, frame_num, max_fn)
                            diff = pred - pn
                            if diff > 0:
                                bw.ue(0); bw.ue(diff - 1)
                            elif diff < 0:
                                bw.ue(1); bw.ue(-diff - 1)
                            else:
                                bw.ue(0); bw.ue(max_fn - 1)      # a full turn back to the same picture number
                            pred = pn
                        else:
                            bw.ue(2); bw.ue(f["lt"])
                        # the chosen picture moves to position idx, the rest shifts back
                        cur_list = cur_list[:idx] + [f] + [g for g in cur_list[idx:] if g is not f]
                    bw.ue(3)
                else:
                    bw.u(1, 0)
            if is_ref:
                if idr:
                    bw.u(1, no_out)
                    bw.u(1, 1 if lt_idr else 0)
                elif mmco:
                    bw.u(1, 1)
                    for op in mmco[0]:
                        for v in op:
                            bw.ue(v)
                    bw.ue(0)
                else:
                    bw.u(1, 0)          # sliding window
            qp_target = r.randint(max(0, pps["pic_init_qp"] - 10), min(51, pps["pic_init_qp"] + 6))
            if knobs["qp_lo"] is not None:
                qp_target = max(knobs["qp_lo"], min(knobs["qp_hi"], qp_target))
            bw.se(qp_target - pps["pic_init_qp"])
            if pps["deblock_ctrl"]:
                idc = r.choice([0, 0, 1, 2])
                bw.ue(idc)
                if idc != 1:
                    bw.se(r.randint(-6, 6)); bw.se(r.randint(-6, 6))
            if fmo["groups"] > 1 and fmo["type"] in (3, 4, 5):
                # Ceil(Log2(PicSizeInMapUnits / SliceGroupChangeRate + 1)) bits, the division not rounded (7.4.3)
                nbits = 0
                while (1 << nbits) * fmo["rate"] < size + fmo["rate"]:
                    nbits += 1
                bw.u(nbits, cycle)
            valid = [i for i in range(min(num_active, len(cur_list))) if not cur_list[i].get("ne")] if is_p else []
            if is_p and not valid:
                valid = None        # no usable entry in this list order: P_Skip / inter would fail -> intra macroblocks only
            pw.write_slice(bw, mbs, is_p, qp_target, num_active, valid)
            return nal((r.choice([1, 2, 3]) if is_ref else 0), 5 if idr else 1, bw.out)

        dropped = None
        if pps["redundant_present"] and len(slices) > 1 and r.random() < 0.6:
            dropped = r.randrange(len(slices))          # this primary slice is "lost"; a redundant one stands in
        for i, mbs in enumerate(slices):
            if i != dropped:
                out += emit_slice(mbs, 0)
        if pps["redundant_present"]:
            extra = []
            if dropped is not None:
                # macroblocks decoded a second time while the picture is still incomplete: the reference keeps the pels of the
                # first decode but filters with the state of the second (h264bsd_macroblock_layer.c:1003-1007,:1108-1111)
                if force.get("redundant_overlap", True):
                    extra = [slices[i] for i in range(len(slices)) if i != dropped and r.random() < 0.4]
                extra.append(slices[dropped])
            extra += [mbs for mbs in slices if r.random() < 0.3]     # after the picture is complete: skipped
            for mbs in extra:
                out += emit_slice(mbs, r.randint(1, 3))

        # decoded reference picture marking (8.2.5)
        if idr:
            refs = []
            max_lt = 0 if lt_idr else None
            refs.append({"fn": 0, "lt": 0 if lt_idr else None})
            prev_ref_fn = 0
        elif is_ref:
            if mmco:
                _, refs, max_lt, cur_lt, had5 = mmco
                refs.append({"fn": 0 if had5 else frame_num, "lt": cur_lt})
                prev_ref_fn = 0 if had5 else frame_num
                if had5:
                    poc_base = 0
                    poc_seen = [0]
            else:
                _sliding_window(refs, nrf, frame_num, max_fn)
                refs.append({"fn": frame_num, "lt": None})
                prev_ref_fn = frame_num
        elif gap:
            prev_ref_fn = (frame_num - 1) % max_fn       # the last gap filler is the "previous reference frame" now (8.2.5.2)
        prev_was_nonref = not is_ref
        after5 = bool(mmco and mmco[4])
        if r.random() < 0.05:
            out += nal(0, 9, bytes([0x10]))          # access unit delimiter
        if r.random() < 0.05:
            out += nal(0, 6, bytes([1, 1, 0, 0x80]))  # SEI (the reference skips it)
    if r.random() < 0.2:
        out += nal(0, 10, b"")                        # end of sequence
    return bytes(out)


# ---------------------------------------------------------------------------------------------- damage
def corrupt_stream(data, seed):
    """A damaged copy of `data` (deterministic in `seed`): flipped bits, a lost slice NAL, a deleted byte range or a burst of
    random bytes -- what makes the decoder mark slices corrupt and conceal (h264bsd_slice_data.c:298-354, h264bsd_conceal.c)."""
    r = random.Random(seed * 7919 + 1)
    e = bytearray(data)
    mode = r.randrange(4)
    if mode == 0:
        for _ in range(r.randint(1, 4)):
            i = r.randrange(60, len(e))
            e[i] ^= 1 << r.randrange(8)
    elif mode == 1:
        starts = [i for i in range(len(e) - 4) if e[i:i + 4] == b"\x00\x00\x00\x01"]
        vcl = [k for k, i in enumerate(starts) if (e[i + 4] & 0x1F) in (1, 5)]
        if len(vcl) > 1:
            k = r.choice(vcl[1:])
            end = starts[k + 1] if k + 1 < len(starts) else len(e)
            del e[starts[k]:end]
    elif mode == 2:
        i = r.randrange(60, len(e))
        j = min(len(e), i + r.randint(1, 60))
        del e[i:j]
    else:
        i = r.randrange(60, len(e))
        for k in range(i, min(len(e), i + r.randint(1, 12))):
            e[k] = r.randrange(256)
    return bytes(e)


def make_damaged_stream(seed):
    """stream `seed % 400` damaged by corrupt_stream(seed).  Without redundant slices: a macroblock that was decoded, served as
    intra neighbour and is then given up because a redundant slice over it turns out corrupt is the one case the tape cannot
    express (deviations)"""
    return corrupt_stream(make_stream(seed % 400, redundant=False), seed)
