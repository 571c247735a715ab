import numpy as np
from scipy.special import erfc, erf
from numpy.polynomial import chebyshev as Ch, polynomial as Pl
TMAX = 5.7
t = np.cos(np.pi * (np.arange(20001) + 0.5) / 20001) * TMAX / 2 + TMAX / 2
q = np.log2(0.5 * erfc(t / np.sqrt(2)))
w = t * 0.5 * erfc(t / np.sqrt(2)) + 1e-3       # weight ~ sensitivity of gelu abs error
def gelu64(x): return 0.5 * x * (1 + erf(x / np.sqrt(2)))
xs = np.concatenate([np.linspace(-8, 8, 2000001), np.random.default_rng(0).standard_normal(1000000) * 2]).astype(np.float32)
def eval32(coef, x):
    x = x.astype(np.float32)
    tt = np.minimum(np.abs(x), np.float32(TMAX)).astype(np.float32)
    r = np.full_like(tt, np.float32(coef[-1]))
    for c in coef[-2::-1]:
        r = (r * tt + np.float32(c)).astype(np.float32)      # fma emulated with a rounding after mul+add (pessimistic)
    e = np.exp2(r.astype(np.float64)).astype(np.float32)
    phi = np.where(x > 0, np.float32(1) - e, e).astype(np.float32)
    return (x * phi).astype(np.float32)
for deg in range(7, 14):
    # weighted least squares in Chebyshev basis on [0,TMAX], then iterate reweighting (Lawson) toward minimax
    u = 2 * t / TMAX - 1
    V = Ch.chebvander(u, deg)
    lw = np.ones_like(t)
    for it in range(40):
        W = w * lw
        c, *_ = np.linalg.lstsq(V * W[:, None], q * W, rcond=None)
        err = np.abs(V @ c - q) * w
        lw = lw * (err / err.max() + 1e-3) ** 0.5
        lw /= lw.max()
    # convert to power basis in t
    pc = Ch.cheb2poly(c)
    # substitute u = 2t/TMAX - 1
    P = np.poly1d([0.0])
    base = np.poly1d([2 / TMAX, -1])
    for k, ck in enumerate(pc):
        P = P + ck * base ** k
    coef = P.coeffs[::-1]
    g = eval32(coef, xs)
    ref = gelu64(xs.astype(np.float64))
    e_abs = np.abs(g - ref).max()
    print(deg, "weighted q err", err.max(), "gelu max abs err (fp32 eval)", e_abs, "rel-l2", np.linalg.norm(g - ref) / np.linalg.norm(ref))
    if deg in (9, 10, 11): print("   coef", [float(np.float32(v)) for v in coef])
# baseline: exact fp32 gelu rounding
g0 = gelu64(xs.astype(np.float64)).astype(np.float32)
print("fp32 rounding of exact:", np.abs(g0 - gelu64(xs.astype(np.float64))).max())
coef9 = [-0.9999985098838806, -1.1511284112930298, -0.4590948820114136, -0.05274663493037224, 0.007349733263254166, -0.00031273943022824824, -0.00013128264981787652, 3.314594505354762e-05, -3.4098220567102544e-06, 1.3804051945953688e-07]
g = eval32(coef9, xs); ref = gelu64(xs.astype(np.float64))
rel = np.abs(g - ref) / np.maximum(np.abs(ref), 1e-3)
print("deg9 max rel err (floor 1e-3):", rel.max(), "at x=", xs[rel.argmax()])
m = np.abs(xs) < 3
print("deg9 max abs err |x|<3:", np.abs(g - ref)[m].max())
g0 = ref.astype(np.float32); rel0 = np.abs(g0 - ref) / np.maximum(np.abs(ref), 1e-3); print("exact-rounded max rel", rel0.max())
