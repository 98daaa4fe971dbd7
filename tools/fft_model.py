"""NumPy model of the per-thread FFT dataflow used by csrc/gss_fft.cuh.

The CUDA kernels cannot be run in the build container (no GPU), so the index
maps - which thread holds which sample, the two shared-memory exchange
layouts, the twiddle tables, the in-thread Hermitian pairing and the thread-0
fix-up - are modelled here one to one (arrays are [thread, register]) and
checked against numpy.fft in tests/test_fft_model.py, together with a
bank-conflict count for every shared-memory access pattern.

Geometry:  N = 64*M complex points per transform (M in {4, 8, 16}),
TPF = 4*M threads per transform, 16 complex points per thread, three passes
radix 8 / M / 8;  L = N/8.

  pass 0 : thread j owns butterflies n' = 2j+e (e=0,1) over x[n' + L*n0]
  middle : DFT-M over n1 for (k0, n2), n' = 8*n1 + n2
  last   : thread j owns butterflies c in {cA=j, cB=L-j} (thread 0: {0, L/2}),
           output Z[c + L*k2]  ->  Z[k] and Z[N-k] sit in the same thread.
"""
from __future__ import annotations

import numpy as np


class Geometry:
    def __init__(self, N):
        assert N in (256, 512, 1024)
        self.N = N
        self.M = N // 64
        self.TPF = 4 * self.M
        self.L = N // 8
        # E0: two planes (re, im) of 8-byte v2 units, [k0][j] -> (x(2j), x(2j+1)); pitch avoids conflicts
        # (csrc/gss_fft.cuh: c.e0[k0*P0 + j] and c.e0[E0_PLANE + k0*P0 + j], 64-bit accesses)
        self.P0 = self.TPF + 4
        self.E0_PLANE_V2 = 8 * self.P0
        # E1: float units, two planes (re, im), [n2][c]
        self.P1 = self.L + 4
        self.E0_FLOATS = 8 * self.P0 * 4
        self.E1_PLANE = 8 * self.P1
        self.E1_FLOATS = 2 * self.E1_PLANE

    # ---- thread -> butterfly maps --------------------------------------
    def mid_butterflies(self, j):
        """list of (k0, n2) handled by thread j in the middle pass, as SIMD
        pairs where possible: returns list of tuples of 1 or 2 butterflies."""
        M = self.M
        if M == 8:
            k0, q = j // 4, j % 4
            return [((k0, 2 * q), (k0, 2 * q + 1))]
        if M == 4:
            k0, q = j // 4, j % 4
            return [((k0, 2 * q), (k0, 2 * q + 1)), ((k0 + 4, 2 * q), (k0 + 4, 2 * q + 1))]
        if M == 16:
            k0, n2 = j // 8, j % 8
            return [((k0, n2),)]
        raise ValueError

    def last_c(self, j):
        return (j, self.L - j) if j else (0, self.L // 2)

    # ---- smem addresses (float index) ----------------------------------
    def e0_addr(self, k0, nprime):
        """float indices (re, im) of element (k0; n')."""
        a2 = k0 * self.P0 + nprime // 2
        e = nprime & 1
        return 2 * a2 + e, 2 * (self.E0_PLANE_V2 + a2) + e

    def e1_addr(self, c, n2):
        f = n2 * self.P1 + c
        return f, f + self.E1_PLANE


def w(N, e):
    return np.exp(-2j * np.pi * (e % N) / N)


def forward(x, N):
    """x: [N] complex.  Returns (ZA, ZB, cA, cB): per-thread last-pass
    registers ZA[j,k2], ZB[j,k2] AFTER the thread-0 fix-up."""
    g = Geometry(N)
    M, TPF, L = g.M, g.TPF, g.L
    E0 = np.zeros(g.E0_FLOATS)
    E1 = np.zeros(g.E1_FLOATS)
    # pass 0
    for j in range(TPF):
        for e in range(2):
            npr = 2 * j + e
            inp = np.array([x[npr + L * n0] for n0 in range(8)])
            a = np.array([sum(inp[n0] * w(8, n0 * k0) for n0 in range(8)) for k0 in range(8)])
            a = a * np.array([w(N, npr * k0) for k0 in range(8)])
            for k0 in range(8):
                r, i = g.e0_addr(k0, npr)
                E0[r], E0[i] = a[k0].real, a[k0].imag
    # middle
    for j in range(TPF):
        for group in g.mid_butterflies(j):
            for (k0, n2) in group:
                inp = []
                for n1 in range(M):
                    r, i = g.e0_addr(k0, 8 * n1 + n2)
                    inp.append(E0[r] + 1j * E0[i])
                b = np.array([sum(inp[n1] * w(M, n1 * k1) for n1 in range(M)) for k1 in range(M)])
                b = b * np.array([w(L, n2 * k1) for k1 in range(M)])
                for k1 in range(M):
                    r, i = g.e1_addr(k0 + 8 * k1, n2)
                    E1[r], E1[i] = b[k1].real, b[k1].imag
    # last
    ZA = np.zeros((TPF, 8), complex)
    ZB = np.zeros((TPF, 8), complex)
    for j in range(TPF):
        cA, cB = g.last_c(j)
        for c, Z in ((cA, ZA), (cB, ZB)):
            inp = []
            for n2 in range(8):
                r, i = g.e1_addr(c, n2)
                inp.append(E1[r] + 1j * E1[i])
            Z[j] = [sum(inp[n2] * w(8, n2 * k2) for n2 in range(8)) for k2 in range(8)]
    fixup_thread0(ZA, ZB)
    return ZA, ZB


def fixup_thread0(ZA, ZB):
    """A''[4..7] = B[4..7];  B''[4..7] = [A5, A6, A7, A4]   (thread 0 only)."""
    a = ZA[0].copy()
    b = ZB[0].copy()
    ZA[0, 4:8] = b[4:8]
    ZB[0, 4:8] = [a[5], a[6], a[7], a[4]]


def unfix_thread0(ZA, ZB):
    a = ZA[0].copy()
    b = ZB[0].copy()
    ZB[0, 4:8] = a[4:8]
    ZA[0, 4:8] = [b[7], b[4], b[5], b[6]]


def pair_k(N, j, group, i):
    """frequency index k (<= N/2) of pair (group, i) in thread j, i in 0..3.
    group 0: (P, Q) = (ZA[i], ZB[7-i]);  group 1: (P, Q) = (ZB[i], ZA[7-i]).
    Thread 0, group 0, i = 0 is the special (Z[0], Z[N/2]) pair."""
    g = Geometry(N)
    cA, cB = g.last_c(j)
    return (cA if group == 0 else cB) + g.L * i


def separate(ZA, ZB, N):
    """two-for-one split: frames a (real part) and b (imag part).
    Returns Xa[k], Xb[k] for k in 0..N/2 assembled from the per-thread pairs."""
    g = Geometry(N)
    Xa = np.zeros(N // 2 + 1, complex)
    Xb = np.zeros(N // 2 + 1, complex)
    seen = np.zeros(N // 2 + 1, int)
    for j in range(g.TPF):
        for grp in range(2):
            for i in range(4):
                P, Q = (ZA[j, i], ZB[j, 7 - i]) if grp == 0 else (ZB[j, i], ZA[j, 7 - i])
                k = pair_k(N, j, grp, i)
                if j == 0 and grp == 0 and i == 0:
                    Xa[0], Xb[0] = P.real, P.imag
                    Xa[N // 2], Xb[N // 2] = Q.real, Q.imag
                    seen[0] += 1
                    seen[N // 2] += 1
                    continue
                Xa[k] = 0.5 * (P + np.conj(Q))
                Xb[k] = -0.5j * (P - np.conj(Q))
                seen[k] += 1
    assert np.all(seen == 1), seen
    return Xa, Xb


def hermitian_pack(Ya, Yb, N):
    """inverse of separate(): per-thread (ZA, ZB) holding Z = Ya + i*Yb."""
    g = Geometry(N)
    ZA = np.zeros((g.TPF, 8), complex)
    ZB = np.zeros((g.TPF, 8), complex)
    for j in range(g.TPF):
        for grp in range(2):
            for i in range(4):
                k = pair_k(N, j, grp, i)
                if j == 0 and grp == 0 and i == 0:
                    P = Ya[0].real + 1j * Yb[0].real
                    Q = Ya[N // 2].real + 1j * Yb[N // 2].real
                else:
                    P = Ya[k] + 1j * Yb[k]
                    Q = np.conj(Ya[k]) + 1j * np.conj(Yb[k])
                if grp == 0:
                    ZA[j, i], ZB[j, 7 - i] = P, Q
                else:
                    ZB[j, i], ZA[j, 7 - i] = P, Q
    return ZA, ZB


def inverse(ZA, ZB, N):
    """transposed network; input in the (fixed-up) last-pass layout, output
    y[n] = sum_k Z[k] e^{+2 pi i nk/N} as per-thread registers y[j, e, n0]
    (sample n = 2j+e + L*n0)."""
    g = Geometry(N)
    M, TPF, L = g.M, g.TPF, g.L
    ZA = ZA.copy()
    ZB = ZB.copy()
    unfix_thread0(ZA, ZB)
    E0 = np.zeros(g.E0_FLOATS)
    E1 = np.zeros(g.E1_FLOATS)
    for j in range(TPF):
        cA, cB = g.last_c(j)
        for c, Z in ((cA, ZA), (cB, ZB)):
            b = [sum(Z[j, k2] * np.conj(w(8, n2 * k2)) for k2 in range(8)) for n2 in range(8)]
            for n2 in range(8):
                r, i = g.e1_addr(c, n2)
                E1[r], E1[i] = b[n2].real, b[n2].imag
    for j in range(TPF):
        for group in g.mid_butterflies(j):
            for (k0, n2) in group:
                inp = []
                for k1 in range(M):
                    r, i = g.e1_addr(k0 + 8 * k1, n2)
                    inp.append((E1[r] + 1j * E1[i]) * np.conj(w(L, n2 * k1)))
                a = [sum(inp[k1] * np.conj(w(M, n1 * k1)) for k1 in range(M)) for n1 in range(M)]
                for n1 in range(M):
                    r, i = g.e0_addr(k0, 8 * n1 + n2)
                    E0[r], E0[i] = a[n1].real, a[n1].imag
    y = np.zeros((TPF, 2, 8), complex)
    for j in range(TPF):
        for e in range(2):
            npr = 2 * j + e
            inp = []
            for k0 in range(8):
                r, i = g.e0_addr(k0, npr)
                inp.append((E0[r] + 1j * E0[i]) * np.conj(w(N, npr * k0)))
            y[j, e] = [sum(inp[k0] * np.conj(w(8, n0 * k0)) for k0 in range(8)) for n0 in range(8)]
    return y


def registers_to_positions(y, N):
    g = Geometry(N)
    out = np.zeros(N, complex)
    for j in range(g.TPF):
        for e in range(2):
            for n0 in range(8):
                out[2 * j + e + g.L * n0] = y[j, e, n0]
    return out


# ---------------------------------------------------------------------------
# bank-conflict accounting: each entry is one warp-level instruction given as a
# list of (thread, float_index, width_in_floats)
# ---------------------------------------------------------------------------
def wavefronts(accesses, width):
    """number of shared-memory wavefronts for one warp instruction.
    width 1: all 32 threads in one phase; width 2: half-warps; width 4: quarter-warps."""
    per_phase = 32 // width
    total = 0
    accesses = sorted(accesses)
    for p in range(0, 32, per_phase):
        banks = {}
        for (t, f) in accesses:
            if p <= t % 32 < p + per_phase:
                for wd in range(width):
                    banks.setdefault((f + wd) % 32, set()).add((f + wd) // 32)
        if banks:
            total += max(len(v) for v in banks.values())
    return total


def conflict_report(N):
    """returns dict name -> (wavefronts, ideal) summed over one transform's
    warp instructions for warp 0 of the team."""
    g = Geometry(N)
    M, TPF, L = g.M, g.TPF, g.L
    rep = {}
    lanes = range(min(32, TPF)) if TPF >= 32 else range(32)   # two teams share a warp when TPF=16

    def team_thread(t):
        return t % TPF, t // TPF      # (j, team) for TPF=16

    def add(name, acc, width):
        wf = wavefronts(acc, width)
        ideal = width if TPF >= 32 or True else width
        a, b = rep.get(name, (0, 0))
        rep[name] = (a + wf, b + width)

    team_off = lambda team: team * (g.E0_FLOATS + g.E1_FLOATS)
    # E0 write (pass 0): one v2 per plane per k0
    for k0 in range(8):
        for part in range(2):
            acc = []
            for t in lanes:
                j, team = team_thread(t) if TPF < 32 else (t, 0)
                acc.append((t, team_off(team) + g.e0_addr(k0, 2 * j)[part]))
            add("E0 write (64-bit)", acc, 2)
    # E0 read (middle)
    if M in (4, 8):
        ngroups = len(g.mid_butterflies(0))
        for gi in range(ngroups):
            for n1 in range(M):
                for part in range(2):
                    acc = []
                    for t in lanes:
                        j, team = team_thread(t) if TPF < 32 else (t, 0)
                        (k0, n2), _ = g.mid_butterflies(j)[gi]
                        acc.append((t, team_off(team) + g.e0_addr(k0, 8 * n1 + n2)[part]))
                    add("E0 read (64-bit)", acc, 2)
    else:
        for n1 in range(M):
            for part in range(2):
                acc = []
                for t in lanes:
                    (k0, n2), = g.mid_butterflies(t)[0]
                    acc.append((t, g.e0_addr(k0, 8 * n1 + n2)[part]))
                add("E0 read (32-bit)", acc, 1)
    # E1 write (middle): scalar
    ngroups = len(g.mid_butterflies(0))
    for gi in range(ngroups):
        nb = len(g.mid_butterflies(0)[gi])
        for bi in range(nb):
            for k1 in range(M):
                for part in range(2):
                    acc = []
                    for t in lanes:
                        j, team = team_thread(t) if TPF < 32 else (t, 0)
                        k0, n2 = g.mid_butterflies(j)[gi][bi]
                        acc.append((t, team_off(team) + g.e1_addr(k0 + 8 * k1, n2)[part]))
                    add("E1 write (32-bit)", acc, 1)
    # E1 read (last): scalar
    for side in range(2):
        for n2 in range(8):
            for part in range(2):
                acc = []
                for t in lanes:
                    j, team = team_thread(t) if TPF < 32 else (t, 0)
                    c = g.last_c(j)[side]
                    acc.append((t, team_off(team) + g.e1_addr(c, n2)[part]))
                add("E1 read (32-bit)", acc, 1)
    return rep


if __name__ == "__main__":
    for N in (256, 512, 1024):
        print(N, conflict_report(N))
