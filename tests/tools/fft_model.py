"""Numpy model of csrc/fft_core.cuh's index algebra (DIF forward / DIT inverse with
16 points per thread, digit-reversed spectrum layout, two-level split).  Used by the CPU
tests to prove the index maps of the CUDA kernels without a GPU."""
import numpy as np

LOGE, E = 4, 16          # default points per thread; set_radix() switches the model


def set_radix(loge):
    global LOGE, E
    LOGE, E = loge, 1 << loge


def brev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def dif_regs(a, rad):
    a = list(a)
    h = rad // 2
    while h >= 1:
        for blk in range(0, rad, 2 * h):
            for i in range(h):
                u, v = a[blk + i], a[blk + i + h]
                a[blk + i] = u + v
                a[blk + i + h] = (u - v) * np.exp(-2j * np.pi * i / (2 * h))
        h //= 2
    return a


def dit_regs(a, rad):
    a = list(a)
    h = 1
    while h <= rad // 2:
        for blk in range(0, rad, 2 * h):
            for i in range(h):
                u, v = a[blk + i], a[blk + i + h] * np.exp(2j * np.pi * i / (2 * h))
                a[blk + i] = u + v
                a[blk + i + h] = u - v
        h *= 2
    return a


class Slot:
    def __init__(self, log2l):
        self.log2l = log2l
        self.L = 1 << log2l
        self.NT = self.L // E
        self.nfull = log2l // LOGE
        self.rem = log2l % LOGE
        self.RR = 1 << self.rem

    def pos_of_reg(self, i, t):
        if self.rem == 0:
            return t * E + brev(i, LOGE)
        if self.rem == 1:       # cross-lane radix-2 (SlotFFT::REM_SHFL): positions P and P^1 sit in threads t and t^1
            return ((t >> 1) << (1 + LOGE)) + (t & 1) + (brev(i, LOGE) << 1)
        return (t + (i // self.RR) * self.NT) * self.RR + brev(i % self.RR, self.rem)

    def freq_of_pos(self, P):
        k, mult = 0, 1
        for s in range(self.nfull):
            k += ((P >> (self.log2l - LOGE * (s + 1))) & (E - 1)) * mult
            mult <<= LOGE
        if self.rem:
            k += (P & (self.RR - 1)) * mult
        return k

    def pos_of_freq(self, k):
        P = 0
        for s in range(self.nfull):
            P |= (k & (E - 1)) << (self.log2l - LOGE * (s + 1))
            k >>= LOGE
        if self.rem:
            P |= k & (self.RR - 1)
        return P

    def layout_of_pos(self, P):
        if self.rem == 0:
            t, c = P >> LOGE, P & (E - 1)
            return brev(c, LOGE) * self.NT + t
        if self.rem == 1:
            g, c, r = P >> (1 + LOGE), (P >> 1) & (E - 1), P & 1
            return brev(c, LOGE) * self.NT + ((g << 1) | r)
        b, c = P >> self.rem, P & (self.RR - 1)
        t, u = b % self.NT, b // self.NT
        return (u * self.RR + brev(c, self.rem)) * self.NT + t

    def forward(self, x):
        """returns regs[t][i] exactly as the kernel holds them after forward()."""
        L, NT = self.L, self.NT
        sm = np.zeros(L, dtype=complex)
        regs = [[x[t + c * NT] for c in range(E)] for t in range(NT)]
        for k in range(self.nfull):
            ls = self.log2l - LOGE * (k + 1)
            s = 1 << ls
            for t in range(NT):
                g, r = t >> ls, t & (s - 1)
                base = (g << (ls + LOGE)) + r
                a = regs[t]
                if k > 0:
                    a = [sm[base + c * s] for c in range(E)]
                a = dif_regs(a, E)
                if ls > 0:
                    for c in range(1, E):
                        a[brev(c, LOGE)] *= np.exp(-2j * np.pi * (r * c) / (E * s))
                regs[t] = a
            last = (k == self.nfull - 1) and self.rem in (0, 1)
            if not last:
                for t in range(NT):
                    g, r = t >> ls, t & (s - 1)
                    base = (g << (ls + LOGE)) + r
                    for c in range(E):
                        sm[base + c * s] = regs[t][brev(c, LOGE)]
        if self.rem == 1:
            regs = self._rem_shuffle(regs)
        elif self.rem:
            RR = self.RR
            for t in range(NT):
                a = [0] * E
                for u in range(E // RR):
                    b = t + u * NT
                    sub = [sm[b * RR + c] for c in range(RR)]
                    a[u * RR:(u + 1) * RR] = dif_regs(sub, RR)
                regs[t] = a
        return regs

    def _rem_shuffle(self, regs):
        """the radix-2 level between threads t and t^1: even keeps mine + partner, odd keeps partner - mine"""
        out = []
        for t in range(self.NT):
            mine, other = regs[t], regs[t ^ 1]
            out.append([(other[i] - mine[i]) if (t & 1) else (mine[i] + other[i]) for i in range(E)])
        return out

    def inverse(self, regs):
        L, NT = self.L, self.NT
        sm = np.zeros(L, dtype=complex)
        regs = [list(r) for r in regs]
        if self.rem == 1:
            regs = self._rem_shuffle(regs)
        elif self.rem:
            RR = self.RR
            for t in range(NT):
                for u in range(E // RR):
                    sub = dit_regs(regs[t][u * RR:(u + 1) * RR], RR)
                    b = t + u * NT
                    for c in range(RR):
                        sm[b * RR + c] = sub[c]
        for k in range(self.nfull - 1, -1, -1):
            ls = self.log2l - LOGE * (k + 1)
            s = 1 << ls
            first = (k == self.nfull - 1) and self.rem in (0, 1)
            for t in range(NT):
                g, r = t >> ls, t & (s - 1)
                base = (g << (ls + LOGE)) + r
                a = regs[t]
                if not first:
                    a = [0] * E
                    for c in range(E):
                        a[brev(c, LOGE)] = sm[base + c * s]
                if ls > 0:
                    for c in range(1, E):
                        a[brev(c, LOGE)] *= np.exp(2j * np.pi * (r * c) / (E * s))
                regs[t] = dit_regs(a, E)
            if k > 0:
                for t in range(NT):
                    g, r = t >> ls, t & (s - 1)
                    base = (g << (ls + LOGE)) + r
                    for c in range(E):
                        sm[base + c * s] = regs[t][c]
        y = np.zeros(L, dtype=complex)
        for t in range(NT):
            for j in range(E):
                y[t + j * NT] = regs[t][j]
        return y

    def spectrum_layout(self, regs):
        out = np.zeros(self.L, dtype=complex)
        for t in range(self.NT):
            for i in range(E):
                out[i * self.NT + t] = regs[t][i]
        return out


def two_level_forward(x, l1, l2):
    """Spectrum in the private layout: rows p (N1 of them) of layout-ordered length-N2 rows."""
    A, B = Slot(l1), Slot(l2)
    N1, N2 = 1 << l1, 1 << l2
    Np = N1 * N2
    scratch = np.zeros((N1, N2), dtype=complex)
    for j2 in range(N2):
        regs = A.forward([x[j1 * N2 + j2] for j1 in range(N1)])
        for t in range(A.NT):
            for i in range(E):
                P = A.pos_of_reg(i, t)
                k1 = A.freq_of_pos(P)
                scratch[P, j2] = regs[t][i] * np.exp(-2j * np.pi * (k1 * j2) / Np)
    spec = np.zeros((N1, N2), dtype=complex)
    for p in range(N1):
        spec[p] = B.spectrum_layout(B.forward(scratch[p]))
    return spec


def unscramble(spec_layout, l1, l2):
    B = Slot(l2)
    if l1 == 0:
        L = 1 << l2
        return np.array([spec_layout[B.layout_of_pos(B.pos_of_freq(k))] for k in range(L)])
    A = Slot(l1)
    N1, N2 = 1 << l1, 1 << l2
    out = np.zeros(N1 * N2, dtype=complex)
    flat = spec_layout.reshape(-1)
    for k in range(N1 * N2):
        k1, k2 = k & (N1 - 1), k >> l1
        p = A.pos_of_freq(k1)
        out[k] = flat[(p << l2) + B.layout_of_pos(B.pos_of_freq(k2))]
    return out
