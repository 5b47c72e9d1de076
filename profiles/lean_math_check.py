import numpy as np, math
from mpmath import mp, mpf
mp.prec = 120
rng = np.random.default_rng(0)
LG = [6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
      1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01]
LN2_HI, LN2_LO = 6.93147180369123816490e-01, 1.90821492927058770002e-10
def log_pos(x):
    x = np.asarray(x, dtype=np.float64)
    bits = x.view(np.int64)
    e = ((bits >> 52) & 0x7ff) - 1023
    m = ((bits & 0x000fffffffffffff) | (1023 << 52)).view(np.float64)
    big = m > 1.4142135623730951
    m = np.where(big, m * 0.5, m); e = e + big
    f = m - 1.0
    s = f / (2.0 + f); z = s * s
    R = LG[6]
    for c in LG[5::-1]: R = R * z + c
    R = R * z
    lm = 2.0 * s + s * R
    return e * LN2_HI + (lm + e * LN2_LO)
def relerr(a, b): return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
x = np.concatenate([rng.random(2000000), 10.0 ** rng.uniform(-12, 3, 2000000), 1.0 + rng.uniform(-1e-3, 1e-3, 100000)])
ref = np.array([float(mp.log(mpf(float(v)))) for v in x[::97]])
print("log relerr vs mp", relerr(log_pos(x[::97]), ref), " np.log:", relerr(np.log(x[::97]), ref))
# exp
INV_LN2 = 1.4426950408889634
EC = [1.0 / math.factorial(k) for k in range(14)]
def exp_r(y):
    k = np.rint(y * INV_LN2)
    r = (y - k * LN2_HI) - k * LN2_LO
    p = EC[13]
    for c in EC[12::-1]: p = p * r + c
    return np.ldexp(p, k.astype(np.int64))
y = rng.uniform(-60, 60, 400000)
ref = np.array([float(mp.exp(mpf(float(v)))) for v in y[::41]])
print("exp relerr vs mp", relerr(exp_r(y[::41]), ref), " np.exp:", relerr(np.exp(y[::41]), ref))
# sincospi(2u)
SC = [(-1) ** k * math.pi ** (2 * k + 1) / math.factorial(2 * k + 1) for k in range(10)]   # sin(pi r)/r in r^2
CC = [(-1) ** k * math.pi ** (2 * k) / math.factorial(2 * k) for k in range(11)]           # cos(pi r) in r^2
def sincospi2(u):
    t = 2.0 * u
    j = np.rint(2.0 * t)
    r = t - 0.5 * j
    z = r * r
    sp = SC[8]
    for c in SC[7::-1]: sp = sp * z + c
    sp = sp * r
    cp = CC[9]
    for c in CC[8::-1]: cp = cp * z + c
    q = j.astype(np.int64) & 3
    s = np.where(q == 0, sp, np.where(q == 1, cp, np.where(q == 2, -sp, -cp)))
    c = np.where(q == 0, cp, np.where(q == 1, -sp, np.where(q == 2, -cp, sp)))
    return s, c
u = (rng.integers(0, 2**52, 400000) + 0.5) * 2.0 ** -52
s, c = sincospi2(u)
uu = u[::41]
rs = np.array([float(mp.sin(2 * mp.pi * mpf(float(v)))) for v in uu]); rc = np.array([float(mp.cos(2 * mp.pi * mpf(float(v)))) for v in uu])
print("sincos abs err vs mp", np.max(np.abs(s[::41] - rs)), np.max(np.abs(c[::41] - rc)), " numpy:", np.max(np.abs(np.sin(2*np.pi*uu) - rs)))
print("unit circle", np.max(np.abs(s*s + c*c - 1)))
# sqrt via rsqrt newton from float seed
def sqrt_pos(x):
    y = (1.0 / np.sqrt(x.astype(np.float32))).astype(np.float64)     # ~24-bit seed
    y = y * (1.5 - 0.5 * x * y * y)
    y = y * (1.5 - 0.5 * x * y * y)
    g = x * y
    return g + (x - g * g) * (0.5 * y)
x = np.concatenate([10.0 ** rng.uniform(-12, 4, 1000000), rng.uniform(0, 80, 1000000)])
print("sqrt relerr", relerr(sqrt_pos(x), np.sqrt(x)))
