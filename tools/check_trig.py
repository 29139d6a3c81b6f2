"""Accuracy of the table trig used by the FP32 IK kernels (csrc/pnp_common.cuh, Trig<float> / TrigV: first order on an
8192-entry sine table) and, for comparison, of the first session's 1024-entry second-order scheme,
emulated in NumPy with exact FMA semantics (products of two floats are exact in float64).
Prints the max abs / rms error against float64 sin/cos over several ranges."""
import numpy as np

f32 = np.float32
N = 1024
DELTA = 2 * np.pi / N
D_HI = f32(DELTA)
D_LO = f32(DELTA - float(D_HI))
INV = f32(N / (2 * np.pi))
K = np.arange(N)
TAB_S, TAB_C = np.sin(K * DELTA).astype(f32), np.cos(K * DELTA).astype(f32)


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def sincos_tab(x):
    x = x.astype(f32)
    full = lambda v: np.full_like(x, f32(v))  # noqa: E731
    mag = f32(12582912.0)
    t = fma(x, full(INV), full(mag))
    ji = t.view(np.int32)
    k = (t - mag).astype(f32)
    r = fma(k, full(-D_HI), x)
    r = fma(k, full(-D_LO), r)
    idx = ji & (N - 1)
    sk, ck = TAB_S[idx], TAB_C[idx]
    h = ((f32(0.5) * r).astype(f32) * r).astype(f32)
    return fma(ck, r, fma(-sk, h, sk)), fma(-sk, r, fma(-ck, h, ck))


# ---- TrigV of the value-type kernels: first order on an 8192-entry table (one sin table, cos = +2048 entries)
NV = 8192
DV = 2 * np.pi / NV
DV_HI = f32(DV)
DV_LO = f32(DV - float(DV_HI))
INVV = f32(NV / (2 * np.pi))
TABV = np.sin(np.arange(NV + NV // 4) * DV).astype(f32)


TWO_TERM = False


def sincos_tabv(x):
    x = x.astype(f32)
    full = lambda v: np.full_like(x, f32(v))  # noqa: E731
    mag = f32(12582912.0)
    t = fma(x, full(INVV), full(mag))
    ji = t.view(np.int32) & (NV - 1)
    k = (t - mag).astype(f32)
    r = fma(k, full(-DV_HI), x)
    if TWO_TERM:  # rounds 1 and 2 of the build: a second Cody-Waite term (one more packed instruction per joint and pass)
        r = fma(k, full(-DV_LO), r)
    sk, ck = TABV[ji], TABV[ji + NV // 4]
    return fma(ck, r, sk), fma(-sk, r, ck)


def report(name, fn, lims):
    rng = np.random.default_rng(0)
    for lim in lims:
        x = rng.uniform(-lim, lim, 2_000_000).astype(f32)
        s, c = fn(x)
        xs = x.astype(np.float64)
        es, ec = np.abs(s - np.sin(xs)), np.abs(c - np.cos(xs))
        print(f"{name} |x| < {lim:8g}: max abs err sin {es.max():.3e} cos {ec.max():.3e}  rms {np.sqrt(np.mean(es ** 2)):.2e}")


if __name__ == "__main__":
    print(f"first session: INV={INV:.9g} D_HI={D_HI:.17g} D_LO={D_LO:.17g}")
    report("first session (1024, 2nd order)", sincos_tab, (4.0, 100.0, 1e4, 2.5e4))
    print(f"Trig<float> / TrigV: INV={INVV:.17g} D_HI={DV_HI:.17g} -D_LO={-DV_LO:.17g}")
    report("Trig<float>/TrigV (8192, 1st order, one-term reduction)", sincos_tabv, (3.8, 6.3, 100.0, 1e3))
    TWO_TERM = True
    report("same with the second reduction term (until round 2)", sincos_tabv, (3.8, 100.0, 1e3, 3.2e3))
