"""Development probe run under gpurun: first-contact checks of every kernel + rough timings."""
import os, sys, time, traceback, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from zsc_b200 import Engine, datagen, zsc
from refimpl import ref, model_deflate, have_ref

def step(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"[ok ] {name}: {r} ({time.time()-t:.2f}s)", flush=True)
    except Exception as e:
        print(f"[ERR] {name}: {e!r}", flush=True)
        traceback.print_exc()

E = Engine(raw_bytes=600 << 20, comp_bytes=700 << 20, deflate_batch_max=600 << 20)
print(E.L.zscgpu_build_info().decode())
R = ref() if have_ref() else None
N = 4 << 20
xm = datagen.fill(N, 1, 0)
xt = datagen.fill(N, 1000, 1, piece=262144)

def chk_adler():
    E.upload(0, 0, xm)
    out = []
    for n in (0, 1, 15, 16, 17, 5552, 65536, 65537, 1000003, N):
        a = E.adler32(0, n); c = E.crc32(0, n)
        ea = zlib.adler32(xm[:n].tobytes()); ec = zlib.crc32(xm[:n].tobytes())
        out.append((n, a == ea, c == ec))
        if a != ea or c != ec: print("   mismatch", n, hex(a), hex(ea), hex(c), hex(ec))
    # unaligned start
    a = E.adler32(3, 100001); ea = zlib.adler32(xm[3:100004].tobytes())
    c = E.crc32(3, 100001); ec = zlib.crc32(xm[3:100004].tobytes())
    out.append(("unal", a == ea, c == ec))
    return out
step("checksums", chk_adler)

def chk_deflate(x, level, mbl=262144, strategy=0, name=""):
    def f():
        E.upload(0, 0, x)
        cap = len(x) + len(x) // 8 + 4096
        st = Engine.make_streams([0], [len(x)], [0], [cap])
        res = E.deflate(st, mbl, level, strategy)
        r = res[0]
        comp = E.download(1, 0, r.produced)
        info = dict(ret=r.ret, produced=r.produced, adler_ok=(r.check == zlib.adler32(x.tobytes())))
        try:
            back = np.frombuffer(zlib.decompress(comp.tobytes()), dtype=np.uint8)
            info["zlib_ok"] = bool(len(back) == len(x) and (back == x).all())
        except Exception as e:
            info["zlib_ok"] = repr(e)
        model, msyms = model_deflate(x, mbl, level, strategy)
        info["model_size"] = len(model)
        info["bytes_eq_model"] = bool(len(model) == len(comp) and (model == comp).all())
        if not info["bytes_eq_model"]:
            gs, n0 = E.symbols(0, 300000)
            nm = min(len(gs), len(msyms))
            diff = np.nonzero(gs[:nm] != msyms[:nm])[0]
            info["chunk0_nsym"] = n0
            info["first_sym_diff"] = int(diff[0]) if len(diff) else None
            if len(model) and len(comp):
                m = min(len(model), len(comp)); d = np.nonzero(model[:m] != comp[:m])[0]
                info["first_byte_diff"] = int(d[0]) if len(d) else None
        if R is not None:
            rr, out, used = R.uncompress(comp, len(x))
            info["ref_inflate"] = (rr, bool(len(out) == len(x) and (out == x).all()), used == len(comp))
            rc, refc = R.compress(x, mbl, level, strategy=strategy)
            info["ref_size"] = len(refc)
        return info
    step(f"deflate {name} L{level} s{strategy} mbl{mbl}", f)

chk_deflate(xm[:100000], 1, name="mixed100k")
chk_deflate(xm, 1, name="mixed")
chk_deflate(xt, 1, name="telem")
chk_deflate(xm, 6, name="mixed")
chk_deflate(xt, 9, name="telem")
chk_deflate(xm, 1, mbl=100000, name="mixed")
chk_deflate(xm, 6, mbl=1 << 30, name="mixed-onesection")
chk_deflate(xm, 6, strategy=2, name="huff")
chk_deflate(xm, 6, strategy=3, name="rle")
chk_deflate(xm, 6, strategy=4, name="fixed")
chk_deflate(xm, 0, name="stored")
chk_deflate(datagen.random_bytes(1 << 20), 6, name="random")
chk_deflate(np.zeros(1 << 20, np.uint8), 6, name="zeros")
chk_deflate(xm[:0], 6, name="empty")
chk_deflate(xm[:1], 6, name="one")

def chk_inflate():
    out = []
    srcs = [(xm, 1), (xt, 6), (xm, 9), (xm[:100000], 6), (xm[:0], 6), (datagen.random_bytes(300000), 6)]
    comps = []
    for x, lvl in srcs:
        if R is not None: rr, c = R.compress(x, 262144, lvl)
        else: c = np.frombuffer(zlib.compress(x.tobytes(), lvl), dtype=np.uint8)
        comps.append(c)
    coff, roff, offs = 0, 0, []
    for (x, lvl), c in zip(srcs, comps):
        E.upload(1, coff, c)
        offs.append((roff, len(x), coff, len(c)))
        coff += (len(c) + 63) & ~63; roff += (len(x) + 63) & ~63
    st = Engine.make_streams([o[0] for o in offs], [o[1] for o in offs], [o[2] for o in offs], [o[3] for o in offs])
    res = E.inflate(st, 1)
    for (x, lvl), o, r in zip(srcs, offs, res):
        got = E.download(0, o[0], r.produced)
        out.append((r.ret, r.produced == len(x), r.consumed == o[3], bool(len(got) == len(x) and (got == x).all()),
                    r.check == zlib.adler32(x.tobytes())))
    return out
step("inflate batch", chk_inflate)

def chk_api():
    Z = zsc()
    r, c = Z.compress(xm[:500000], 100000, 6)
    r2, back, used = Z.uncompress(c, 500000)
    return dict(cret=r, clen=len(c), uret=r2, ok=bool((back == xm[:500000]).all()), used=used == len(c),
                adler=Z.adler32(xm[:1000]) == zlib.adler32(xm[:1000].tobytes()),
                crc=Z.crc32(xm[:1000]) == zlib.crc32(xm[:1000].tobytes()))
step("zsc_pub api", chk_api)

def timing(level, x, mbl=262144, reps=3):
    def f():
        E.upload(0, 0, x)
        st = Engine.make_streams([0], [len(x)], [0], [len(x) + len(x) // 8 + 4096])
        E.deflate_enqueue(st, mbl, level); res = E.fetch(1)
        ts = []
        for _ in range(reps):
            E.event(0); E.relaunch(); E.event(1); E.sync(); ts.append(E.elapsed_ms(0, 1))
        return dict(ret=res[0].ret, produced=res[0].produced, ms=ts, GBps=len(x) / 1e6 / min(ts))
    step(f"time deflate L{level} {len(x)>>20} MiB", f)

big = datagen.fill(256 << 20, 1, 0)
timing(1, big)
timing(6, big[:64 << 20])
timing(9, big[:64 << 20], reps=1)

def time_inflate():
    x = big[:256 << 20]
    nst = len(x) // 262144
    E.upload(0, 0, x)
    st = Engine.make_streams([i * 262144 for i in range(nst)], [262144] * nst, [i * 300000 for i in range(nst)], [300000] * nst)
    res = E.deflate(st, 262144, 6)
    bad = sum(1 for r in res if r.ret != 0)
    st2 = Engine.make_streams([i * 262144 for i in range(nst)], [262144] * nst, [i * 300000 for i in range(nst)], [r.produced for r in res])
    E.inflate_enqueue(st2, 1); res2 = E.fetch(nst)
    bad2 = sum(1 for r in res2 if r.ret != 0 or r.produced != 262144)
    back = E.download(0, 0, len(x)); same = bool((back == x).all())
    ts = []
    for _ in range(3):
        E.event(0); E.relaunch(); E.event(1); E.sync(); ts.append(E.elapsed_ms(0, 1))
    return dict(bad_deflate=bad, bad_inflate=bad2, roundtrip=same, ms=ts, GBps=len(x) / 1e6 / min(ts))
step("time inflate 1024 streams", time_inflate)

def time_sums():
    n = 512 << 20
    E.upload(0, 0, big); E.upload(0, 256 << 20, big)
    out = {}
    for name, fn in (("adler", E.L.zscgpu_adler32_enqueue), ("crc", E.L.zscgpu_crc32_enqueue)):
        fn(E.h, 0, n); E.sync()
        ts = []
        for _ in range(3):
            E.event(0); fn(E.h, 0, n); E.event(1); E.sync(); ts.append(E.elapsed_ms(0, 1))
        out[name] = n / 1e6 / min(ts)
    return out
step("time checksums GB/s", time_sums)
E.close()
print("probe done")
