"""The block stage of deflate_huff.cu builds its small trees, sorts its keys and merges its large tree in warp / lane forms
(zb_lengths_warp, zb_sort_keys, zs_merge_kernel).  These tests restate those forms lane by lane — one Python list per
register, an index into it per shuffle — and pin them on the serial code of huff_build.h (zh_lengths, compiled for the
host in tests/libzsc_cpuharness.so), which is itself pinned on the reference (trees.c build_tree / gen_bitlen,
reference src/trees.c:420-507, :595-680) by the stream tests.  The GPU tests compare the kernels' streams with that
serial model bit for bit; this file keeps the equivalence argument runnable without a GPU."""
import ctypes as C
import random

import numpy as np
import pytest

import refimpl

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)


def _har():
    H = refimpl.harness()
    H.h_zh_lengths.argtypes = [u32p, C.c_int, C.c_int, u8p]
    H.h_zh_lengths.restype = C.c_int
    H.h_zh_leaf_depths.argtypes = [u32p, C.c_int, u8p]
    H.h_zh_leaf_depths.restype = C.c_int
    return H


def serial_lengths(freq, n, maxbits):
    f = np.zeros(512, np.uint32)
    f[:n] = freq[:n]
    out = np.zeros(512, np.uint8)
    mc = _har().h_zh_lengths(f.ctypes.data_as(u32p), n, maxbits, out.ctypes.data_as(u8p))
    return [int(v) for v in out[:n]], mc


def warp_lengths(freq_in, n, maxbits):
    """zb_lengths_warp (deflate_huff.cu), lane = symbol"""
    L = 32
    f = [int(freq_in[l]) if l < n else 0 for l in range(L)]
    usedm = sum(1 << l for l in range(L) if f[l])
    m = bin(usedm).count("1")
    max_code = usedm.bit_length() - 1 if usedm else -1
    while m < 2:
        if max_code < 2:
            max_code += 1
            node = max_code
        else:
            node = 0
        f[node] = 1
        usedm |= 1 << node
        m += 1
    key = [(f[l] << 9) | l for l in range(L)]
    rank = [sum(1 for j in range(n) if (usedm >> j) & 1 and key[j] < key[l]) for l in range(L)]
    sw = [0] * 32
    for l in range(L):
        if f[l]:
            sw[rank[l]] = key[l]
    skey = [sw[l] if l < m else 0 for l in range(L)]
    wl = [k >> 9 for k in skey]
    wi = [0] * L
    pl = [0] * L
    pi = [0] * L
    a = b = e = 0
    for _ in range(m - 1):
        la, ib = wl[a & 31], wi[b & 31]
        if a < m and (b >= e or la <= ib):
            pl[a] = e; s = la; a += 1
        else:
            pi[b] = e; s = ib; b += 1
        la, ib = wl[a & 31], wi[b & 31]
        if a < m and (b >= e or la <= ib):
            pl[a] = e; s += la; a += 1
        else:
            pi[b] = e; s += ib; b += 1
        wi[e] = s
        e += 1
    di = [0] * L
    for i in range(e - 2, -1, -1):
        di[i] = di[pi[i]] + 1
    dl = [di[pl[l]] + 1 for l in range(L)]
    isleaf = [l < m for l in range(L)]
    ov = [isleaf[l] and dl[l] > maxbits for l in range(L)]
    dl = [maxbits if ov[l] else dl[l] for l in range(L)]
    if any(ov):
        cnt = [0] * L
        for bits in range(1, maxbits + 1):
            cnt[bits] = sum(1 for l in range(L) if isleaf[l] and dl[l] == bits)
        excess = sum(cnt[l] << (maxbits - l) for l in range(1, maxbits + 1)) - (1 << maxbits)
        while excess > 0:
            bits = max(l for l in range(1, maxbits) if cnt[l])
            cnt[bits] -= 1; cnt[bits + 1] += 2; cnt[maxbits] -= 1
            excess -= 1
        dnew = [0] * L
        run = 0
        for bits in range(maxbits, 0, -1):
            for l in range(L):
                if run <= l < run + cnt[bits]:
                    dnew[l] = bits
            run += cnt[bits]
        dl = dnew
    out = [0] * 32
    for l in range(L):
        if isleaf[l]:
            out[skey[l] & 0x1F] = dl[l]
    return out[:n], max_code


def _histogram(rng, n):
    kind = rng.random()
    if kind < 0.3:
        return [rng.choice([0, 0, 1, 2, 5, 100, 3000]) for _ in range(n)]
    if kind < 0.6:
        return [int(2 ** rng.uniform(0, 12)) * rng.choice([0, 1, 1]) for _ in range(n)]
    if kind < 0.8:      # Fibonacci-like: the deepest trees, lengths beyond maxbits
        a, b, out = 1, 1, []
        for _ in range(n):
            out.append(a); a, b = b, a + b
        rng.shuffle(out)
        return [min(v, 8000) for v in out]
    return [rng.randint(0, 3) * rng.choice([0, 1]) for _ in range(n)]


@pytest.mark.parametrize("n,maxbits", [(30, 15), (19, 7)])
def test_warp_form_of_the_small_trees_equals_zh_lengths(n, maxbits):
    rng = random.Random(n)
    cases = [[0] * n, [0] * (n - 1) + [7], [5] + [0] * (n - 1), [0, 3] + [0] * (n - 2), [1] * n]
    cases += [_histogram(rng, n) for _ in range(1500)]
    for fr in cases:
        assert warp_lengths(fr, n, maxbits) == serial_lengths(fr, n, maxbits), fr


def sort_network(keys, m, ept):
    """zb_sort_keys<EPT>: thread t of warp w holds the elements w * 32 * EPT + q * 32 + lane"""
    N = 128 * ept
    x = [keys[i] if i < m else 0xFFFFFFFF for i in range(N)]
    kk = 2
    while kk <= N:
        j = kk >> 1
        while j > 0:
            nx = list(x)
            if j >= 32 * ept or j < 32:          # through shared memory / by shuffle: the partner is element i ^ j
                for i in range(N):
                    y = x[i ^ j]
                    nx[i] = min(x[i], y) if (((i & j) == 0) == ((i & kk) == 0)) else max(x[i], y)
            else:                                  # in another register of the same thread
                for t in range(128):
                    base = (t >> 5) * 32 * ept + (t & 31)
                    for q in range(ept):
                        if (q & (j >> 5)) == 0:
                            iq, ir = base + 32 * q, base + 32 * (q | (j >> 5))
                            lo, hi = min(x[iq], x[ir]), max(x[iq], x[ir])
                            up = (iq & kk) == 0
                            nx[iq], nx[ir] = (lo, hi) if up else (hi, lo)
            x = nx
            j >>= 1
        kk <<= 1
    return x


@pytest.mark.parametrize("ept,mmax", [(1, 128), (2, 256), (4, 286)])
def test_register_sort_network_sorts(ept, mmax):
    rng = random.Random(ept)
    for m in [2, 3, mmax - 1, mmax] + [rng.randint(2, mmax) for _ in range(12)]:
        keys = rng.sample(range(1, 1 << 22), m)
        out = sort_network(keys, m, ept)
        assert out[:m] == sorted(keys)
        assert all(v == 0xFFFFFFFF for v in out[m:])


def merge_depths(weights):
    """zs_merge_kernel: two queues, the parent's number written to the slot a node leaves, slots of the internal nodes turned
    into depths, leaf depths out"""
    INF = 0xFFFFFFFF
    m = len(weights)
    wl = list(weights) + [0xFFFF] * 4
    w = [0] * (m + 4)
    a = b = e = 0
    la, ib = wl[0], INF
    for _ in range(m - 1):
        if a < m and la <= ib:
            s = la; wl[a] = e; a += 1; la = wl[a] if a < m else INF
        else:
            s = ib; w[b] = e; b += 1; ib = w[b] if b < e else INF
        if a < m and la <= ib:
            s += la; wl[a] = e; a += 1; la = wl[a] if a < m else INF
        else:
            s += ib; w[b] = e; b += 1; ib = w[b] if b < e else INF
        w[e] = s & 0xFFFF
        if b == e:
            ib = s
        e += 1
    w[e - 1] = 0
    for i in range(e - 2, -1, -1):
        w[i] = w[w[i]] + 1
    return [min(255, w[wl[i]] + 1) for i in range(m)]


def test_merge_with_links_in_the_queue_slots_gives_the_serial_leaf_depths():
    rng = random.Random(7)
    H = _har()
    for t in range(400):
        n = 286
        if t % 3 == 0:
            fr = [rng.choice([0, 1, 1, 2, 3, 9, 40, 700]) for _ in range(n)]
        elif t % 3 == 1:
            fr = [int(2 ** rng.uniform(0, rng.choice([3, 8, 11]))) * rng.choice([0, 1, 1, 1]) for _ in range(n)]
        else:
            fr = [rng.randint(0, 2) for _ in range(n)]
        while sum(fr) > 8193:                      # a block holds at most 8192 symbols + the end-of-block symbol
            fr[fr.index(max(fr))] //= 2
        f = np.zeros(512, np.uint32); f[:n] = fr
        out = np.zeros(512, np.uint8)
        m = H.h_zh_leaf_depths(f.ctypes.data_as(u32p), n, out.ctypes.data_as(u8p))
        keys = sorted((v << 9) | i for i, v in enumerate(fr) if v)
        if len(keys) < 2:
            continue                               # (dummy symbols: covered by the small-tree test's rule)
        assert m == len(keys)
        assert merge_depths([k >> 9 for k in keys]) == [int(v) for v in out[:m]]
