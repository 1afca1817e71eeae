#!/usr/bin/env python3
"""Generates (and checks) the per-lane routing words of the warp-wide Keccak-f[1600] in csrc/dev_transcript.cuh.

Lane l = x + 5 y (l < 25) holds state word A[x][y]; lanes 25..31 route to themselves.
  A[l]: the four other lanes of column x (4 x 5 bits) | lane holding column x-1 (5 bits) | lane holding column x+1 (5 bits)
  B[l]: rho offset of (x, y) (6 bits) | source lanes, before pi, of b[x], b[x+1], b[x+2] of this lane's row (3 x 5 bits)
The script replays the lane algorithm in Python against a textbook Keccak-f[1600] and checks that against
hashlib.sha3_256(b"") (same permutation, different padding byte)."""
import hashlib
import random

RHO = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]  # RHO[x][y]
RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808a, 0x8000000080008000, 0x000000000000808b, 0x0000000080000001,
      0x8000000080008081, 0x8000000000008009, 0x000000000000008a, 0x0000000000000088, 0x0000000080008009, 0x000000008000000a,
      0x000000008000808b, 0x800000000000008b, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
      0x000000000000800a, 0x800000008000000a, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
M = (1 << 64) - 1


def pi_src(l):
    """lane whose rotated value lands on lane l: (x', y') = (y, 2x + 3y)  =>  x = 3 y' + x', y = x'"""
    xp, yp = l % 5, l // 5
    return (3 * yp + xp) % 5 + 5 * xp


def tables():
    A, B = [], []
    for l in range(32):
        if l < 25:
            x, y = l % 5, l // 5
            s = [(l + 5 * k) % 25 for k in (1, 2, 3, 4)]
            dm1, dp1 = 5 * y + (x + 4) % 5, 5 * y + (x + 1) % 5
            A.append(s[0] | s[1] << 5 | s[2] << 10 | s[3] << 15 | dm1 << 20 | dp1 << 25)
            n1, n2 = 5 * y + (x + 1) % 5, 5 * y + (x + 2) % 5
            B.append(RHO[x][y] | pi_src(l) << 6 | pi_src(n1) << 11 | pi_src(n2) << 16)
        else:
            A.append(l | l << 5 | l << 10 | l << 15 | l << 20 | l << 25)
            B.append(l << 6 | l << 11 | l << 16)
    return A, B


def rol(v, n):
    n %= 64
    return ((v << n) | (v >> (64 - n))) & M if n else v


def textbook(a):
    a = list(a)
    for r in range(24):
        C = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        D = [C[(x + 4) % 5] ^ rol(C[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ D[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], RHO[x][y])
        a = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & M & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] ^= RC[r]
    return a


def lanes(a, A, B):
    v = list(a) + [0] * 7
    for r in range(24):
        c = [v[l] ^ v[A[l] & 31] ^ v[(A[l] >> 5) & 31] ^ v[(A[l] >> 10) & 31] ^ v[(A[l] >> 15) & 31] for l in range(32)]
        d = [c[(A[l] >> 20) & 31] ^ rol(c[(A[l] >> 25) & 31], 1) for l in range(32)]
        t = [rol(v[l] ^ d[l], B[l] & 63) for l in range(32)]
        v = [t[(B[l] >> 6) & 31] ^ ((~t[(B[l] >> 11) & 31]) & M & t[(B[l] >> 16) & 31]) for l in range(32)]
        v[0] ^= RC[r]
    return v[:25]


if __name__ == "__main__":
    A, B = tables()
    s = [0] * 25
    s[0] ^= 0x06
    s[16] ^= 0x80 << 56
    assert b"".join(x.to_bytes(8, "little") for x in textbook(s)[:4]) == hashlib.sha3_256(b"").digest()
    for _ in range(20):
        a = [random.getrandbits(64) for _ in range(25)]
        assert textbook(a) == lanes(a, A, B)
    print("#define ZK_WK_A_INIT {" + ", ".join("0x%08xu" % v for v in A) + "}")
    print("#define ZK_WK_B_INIT {" + ", ".join("0x%08xu" % v for v in B) + "}")
