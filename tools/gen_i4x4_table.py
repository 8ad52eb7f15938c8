#!/usr/bin/env python
"""Generates gIntra4x4Table of recon_kernel.cuh and checks it against the formulas of clause 8.3.1.2.1-8.3.1.2.9
(the arithmetic of Intra4x4VerticalPrediction ... Intra4x4HorizontalUpPrediction, h264bsd_intra_prediction.c:1493-1830).
Edge array: E[0..3] = L3, L2, L1, L0 ; E[4] = M (corner) ; E[5..12] = A0..A7.
Entry = i0 | i1 << 8 | i2 << 16 | kind << 24 ; kind 0: E[i0], 1: (E[i0]+E[i1]+1)>>1, 2: (E[i0]+2E[i1]+E[i2]+2)>>2, 3: DC."""
import random


def ent(kind, i0, i1=0, i2=0):
    for i in (i0, i1, i2):
        assert 0 <= i <= 12
    return i0 | (i1 << 8) | (i2 << 16) | (kind << 24)


def build():
    tab = [[0] * 16 for _ in range(9)]
    for y in range(4):
        for x in range(4):
            p = y * 4 + x
            tab[0][p] = ent(0, 5 + x)
            tab[1][p] = ent(0, 3 - y)
            tab[2][p] = ent(3, 0)
            tab[3][p] = ent(2, 11, 12, 12) if (x == 3 and y == 3) else ent(2, 5 + x + y, 6 + x + y, 7 + x + y)
            c = 4 + x - y
            tab[4][p] = ent(2, c - 1, c, c + 1)
            z, k = 2 * x - y, x - (y >> 1)
            tab[5][p] = (ent(1, 4 + k, 5 + k) if z >= 0 and z % 2 == 0 else ent(2, 3 + k, 4 + k, 5 + k) if z >= 0
                         else ent(2, 3, 4, 5) if z == -1 else ent(2, 4 - y, 5 - y, 6 - y))
            z, k = 2 * y - x, y - (x >> 1)
            tab[6][p] = (ent(1, 4 - k, 3 - k) if z >= 0 and z % 2 == 0 else ent(2, 5 - k, 4 - k, 3 - k) if z >= 0
                         else ent(2, 3, 4, 5) if z == -1 else ent(2, 4 + x, 3 + x, 2 + x))
            i = x + (y >> 1)
            tab[7][p] = ent(2, 5 + i, 6 + i, 7 + i) if y & 1 else ent(1, 5 + i, 6 + i)
            z, k = x + 2 * y, y + (x >> 1)
            tab[8][p] = (ent(0, 0) if z > 5 else ent(2, 1, 0, 0) if z == 5 else ent(2, 3 - k, 2 - k, 1 - k) if z & 1
                         else ent(1, 3 - k, 2 - k))
    return tab


def spec(mode, x, y, A, L, M):
    a = lambda i: M if i == -1 else A[i]
    l = lambda i: M if i == -1 else L[i]
    if mode == 0: return A[x]
    if mode == 1: return L[y]
    if mode == 3: return (A[6] + 3 * A[7] + 2) >> 2 if (x == 3 and y == 3) else (A[x + y] + 2 * A[x + y + 1] + A[x + y + 2] + 2) >> 2
    if mode == 4:
        if x > y: return (a(x - y - 2) + 2 * a(x - y - 1) + a(x - y) + 2) >> 2
        if x < y: return (l(y - x - 2) + 2 * l(y - x - 1) + l(y - x) + 2) >> 2
        return (A[0] + 2 * M + L[0] + 2) >> 2
    if mode == 5:
        z, k = 2 * x - y, x - (y >> 1)
        if z >= 0 and z % 2 == 0: return (a(k - 1) + a(k) + 1) >> 1
        if z >= 0: return (a(k - 2) + 2 * a(k - 1) + a(k) + 2) >> 2
        if z == -1: return (L[0] + 2 * M + A[0] + 2) >> 2
        return (l(y - 1) + 2 * l(y - 2) + l(y - 3) + 2) >> 2
    if mode == 6:
        z, k = 2 * y - x, y - (x >> 1)
        if z >= 0 and z % 2 == 0: return (l(k - 1) + l(k) + 1) >> 1
        if z >= 0: return (l(k - 2) + 2 * l(k - 1) + l(k) + 2) >> 2
        if z == -1: return (L[0] + 2 * M + A[0] + 2) >> 2
        return (a(x - 1) + 2 * a(x - 2) + a(x - 3) + 2) >> 2
    if mode == 7:
        i = x + (y >> 1)
        return (A[i] + A[i + 1] + 1) >> 1 if y % 2 == 0 else (A[i] + 2 * A[i + 1] + A[i + 2] + 2) >> 2
    z, k = x + 2 * y, y + (x >> 1)
    if z > 5: return L[3]
    if z == 5: return (L[2] + 3 * L[3] + 2) >> 2
    return (L[k] + L[k + 1] + 1) >> 1 if z % 2 == 0 else (L[k] + 2 * L[k + 1] + L[k + 2] + 2) >> 2


if __name__ == "__main__":
    tab = build()
    rnd = random.Random(1)
    for _ in range(500):
        A = [rnd.randrange(256) for _ in range(8)]; L = [rnd.randrange(256) for _ in range(4)]; M = rnd.randrange(256)
        E = [L[3], L[2], L[1], L[0], M] + A
        for mode in (0, 1, 3, 4, 5, 6, 7, 8):
            for p in range(16):
                e = tab[mode][p]
                i0, i1, i2, kd = e & 255, (e >> 8) & 255, (e >> 16) & 255, e >> 24
                v = E[i0] if kd == 0 else (E[i0] + E[i1] + 1) >> 1 if kd == 1 else (E[i0] + 2 * E[i1] + E[i2] + 2) >> 2
                assert v == spec(mode, p & 3, p >> 2, A, L, M), (mode, p)
    for row in tab:
        print("    " + ", ".join("0x%08x" % v for v in row) + ",")
