"""Sharding one logical zsc stream over several engines (one per GPU / rank) — host-side logic only.

Sections are independent (reference src/deflate.c:1240-1252: the hash is cleared at every full flush), so a
stream of S sections splits into contiguous section ranges, one per rank, with no data-path collective.
Each rank compresses its range as RAW deflate (`wrap=0`) with `part` bit 1 set unless it owns the last
section (so its last section ends with a full-flush marker instead of a final block); the host then
concatenates the parts behind a zlib header and appends the combined adler32.
"""
from . import capi


def partition(n_sections, world):
    """Contiguous section ranges [lo, hi) per rank; earlier ranks take the remainder."""
    base, rem = divmod(n_sections, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def byte_range(lo, hi, max_block_len, total_len):
    return min(lo * max_block_len, total_len), min(hi * max_block_len, total_len)


def zlib_header(level, strategy=0, window_bits=15):
    """The two header bytes the reference writes (src/deflate.c:1029-1049)."""
    if level == -1:
        level = 6
    lf = 0 if (strategy >= 2 or level < 2) else (1 if level < 6 else (2 if level == 6 else 3))
    h = ((8 + ((window_bits - 8) << 4)) << 8) | (lf << 6)
    h += 31 - (h % 31)
    return bytes([h >> 8, h & 0xFF])


def adler32_combine(a1, a2, len2):
    return capi.lib().zscgpu_adler32_combine(a1, a2, len2)


def crc32_combine(c1, c2, len2):
    return capi.lib().zscgpu_crc32_combine(c1, c2, len2)


def stitch(parts, adlers, lens, level, strategy=0):
    """parts: raw-deflate payloads in rank order; adlers/lens: adler32 and input length of each part."""
    total = 1
    for a, n in zip(adlers, lens):
        total = adler32_combine(total, a, n) if n else total
    return zlib_header(level, strategy) + b"".join(parts) + total.to_bytes(4, "big")
