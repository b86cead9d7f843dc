"""Bit-exact emulation of Julia (<= 1.6) ``MersenneTwister`` = dSFMT-19937.  Test infrastructure.

The reference draws every random number from Julia's *global* MersenneTwister
(dqmc_framework.jl:152-155 ``Random.seed!(mc.p.seed)``, ``rand(opdim,N,M)``;
dqmc_framework.jl:628-635 ``randuniform``; local_updates.jl:31 ``rand()``), and its tests seed
it explicitly (tests_O3.jl:188 ``Random.seed!(123456789)``).  The generator is the third-party
dSFMT library (Saito & Matsumoto, v2.2.3, bundled with Julia; not under /root/reference); this
file restates its published recursion (dSFMT.c: ``do_recursion``, ``gen_rand_array``,
``dsfmt_chk_init_by_array``, ``initial_mask``, ``period_certification``) and Julia's use of it
(stdlib Random/RNGs.jl: 1002-double cache for scalar ``rand()``; ``rand!(::Array{Float64})``
fills the first ``(n-2)÷2·2`` values directly and the rest through the scalar cache).

Pure-Python integers: meant for the few thousand draws the fixtures need, not for benchmarks.
"""
import struct

import numpy as np

_N = 191
_POS1 = 117
_SL1 = 19
_SR = 12
_MSK1 = 0x000FFAFFFFFFFB3F
_MSK2 = 0x000FFDFFFC90FFFD
_FIX1 = 0x90014964B32F4329
_FIX2 = 0x3B8D12AC548A7C7A
_PCV1 = 0x3D84E1AC0DC82880
_PCV2 = 0x0000000000000001
_LOW = 0x000FFFFFFFFFFFFF
_HIGH = 0x3FF0000000000000
_M64 = 0xFFFFFFFFFFFFFFFF
_M32 = 0xFFFFFFFF
_CACHE = 1002  # MT_CACHE_F


def _bits_to_double(u):
    return struct.unpack("<d", struct.pack("<Q", u))[0]


class JuliaMT:
    """``MersenneTwister(seed)`` with scalar ``rand()`` and array ``rand(dims...)``."""

    def __init__(self, seed):
        self.seed(seed)

    # -- dsfmt_chk_init_by_array ---------------------------------------------------------
    def seed(self, seed):
        key = []
        s = int(seed)
        assert s >= 0
        while True:
            key.append(s & _M32)
            s >>= 32
            if s == 0:
                break
        size = (_N + 1) * 4
        lag = 11
        mid = (size - lag) // 2
        st = [0x8B8B8B8B] * size

        def f1(x):
            return ((x ^ (x >> 27)) * 1664525) & _M32

        def f2(x):
            return ((x ^ (x >> 27)) * 1566083941) & _M32

        klen = len(key)
        count = max(klen + 1, size)
        r = f1(st[0] ^ st[mid % size] ^ st[(size - 1) % size])
        st[mid % size] = (st[mid % size] + r) & _M32
        r = (r + klen) & _M32
        st[(mid + lag) % size] = (st[(mid + lag) % size] + r) & _M32
        st[0] = r
        count -= 1
        i, j = 1, 0
        while j < count and j < klen:
            r = f1(st[i] ^ st[(i + mid) % size] ^ st[(i + size - 1) % size])
            st[(i + mid) % size] = (st[(i + mid) % size] + r) & _M32
            r = (r + key[j] + i) & _M32
            st[(i + mid + lag) % size] = (st[(i + mid + lag) % size] + r) & _M32
            st[i] = r
            i = (i + 1) % size
            j += 1
        while j < count:
            r = f1(st[i] ^ st[(i + mid) % size] ^ st[(i + size - 1) % size])
            st[(i + mid) % size] = (st[(i + mid) % size] + r) & _M32
            r = (r + i) & _M32
            st[(i + mid + lag) % size] = (st[(i + mid + lag) % size] + r) & _M32
            st[i] = r
            i = (i + 1) % size
            j += 1
        for _ in range(size):
            r = f2((st[i] + st[(i + mid) % size] + st[(i + size - 1) % size]) & _M32)
            st[(i + mid) % size] ^= r
            r = (r - i) & _M32
            st[(i + mid + lag) % size] ^= r
            st[i] = r
            i = (i + 1) % size
        # 32-bit limbs -> 64-bit words (little endian), [u0,u1] per 128-bit word
        w = [st[2 * k] | (st[2 * k + 1] << 32) for k in range(2 * (_N + 1))]
        for k in range(2 * _N):  # initial_mask
            w[k] = (w[k] & _LOW) | _HIGH
        # period_certification
        t0 = w[2 * _N] ^ _FIX1
        t1 = w[2 * _N + 1] ^ _FIX2
        inner = (t0 & _PCV1) ^ (t1 & _PCV2)
        sh = 32
        while sh > 0:
            inner ^= inner >> sh
            sh >>= 1
        if (inner & 1) != 1:
            w[2 * _N + 1] ^= 1
        self.status = [(w[2 * k], w[2 * k + 1]) for k in range(_N)]
        self.lung = (w[2 * _N], w[2 * _N + 1])
        self.vals = []
        self.idx = 0  # position in the scalar cache; empty when idx == len(vals)

    # -- the recursion, as a stream of 128-bit words -------------------------------------
    def _gen_words(self, nwords):
        """Advance the generator by ``nwords`` 128-bit words (>= N); return them as (u0,u1)."""
        assert nwords >= _N
        s = list(self.status)
        l0, l1 = self.lung
        out = []
        for i in range(nwords):
            a0, a1 = s[i]
            b0, b1 = s[i + _POS1]
            n0 = ((a0 << _SL1) & _M64) ^ (l1 >> 32) ^ ((l1 << 32) & _M64) ^ b0
            n1 = ((a1 << _SL1) & _M64) ^ (l0 >> 32) ^ ((l0 << 32) & _M64) ^ b1
            l0, l1 = n0, n1
            r = ((l0 >> _SR) ^ (l0 & _MSK1) ^ a0, (l1 >> _SR) ^ (l1 & _MSK2) ^ a1)
            s.append(r)
            out.append(r)
        self.status = s[-_N:]
        self.lung = (l0, l1)
        return out

    def _fill_close1_open2(self, n):
        assert n % 2 == 0 and n >= 2 * _N
        res = []
        for u0, u1 in self._gen_words(n // 2):
            res.append(_bits_to_double(u0))
            res.append(_bits_to_double(u1))
        return res

    # -- Julia front end -----------------------------------------------------------------
    def rand(self):
        """Scalar ``rand()``: pop from the 1002-double cache (refilled close1_open2), minus 1."""
        if self.idx >= len(self.vals):
            self.vals = self._fill_close1_open2(_CACHE)
            self.idx = 0
        v = self.vals[self.idx]
        self.idx += 1
        return v - 1.0

    def rand_array(self, *dims):
        """``rand(dims...)`` :: Array{Float64}; returned in Julia (column-major) index order."""
        n = int(np.prod(dims))
        flat = np.empty(n)
        n2 = (n - 2) // 2 * 2
        if n2 < 2 * _N:
            for k in range(n):
                flat[k] = self.rand()
        else:
            v = self._fill_close1_open2(n2)
            flat[:n2] = np.array(v) - 1.0
            for k in range(n2, n):
                flat[k] = self.rand()
        return flat.reshape(dims[::-1]).T.copy() if len(dims) > 1 else flat
