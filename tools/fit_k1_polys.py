# Near-minimax polynomial fits (Chebyshev interpolation in high precision, then monomial coefficients)
import mpmath as mp, numpy as np
mp.mp.dps = 60

def cheb_fit(f, lo, hi, deg):
    # interpolate f at Chebyshev nodes on [lo,hi]; return monomial coeffs in the variable z
    n = deg + 1
    nodes = [ (lo+hi)/2 + (hi-lo)/2*mp.cos(mp.pi*(2*k+1)/(2*n)) for k in range(n)]
    A = mp.matrix(n, n)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x**j
    b = mp.matrix([f(x) for x in nodes])
    c = mp.lu_solve(A, b)
    return [c[j] for j in range(n)]

def horner_fp64(coefs, z):
    acc = np.full_like(z, float(coefs[-1]))
    for c in coefs[-2::-1]:
        acc = acc * z + float(c)   # numpy: no fma, slightly pessimistic
    return acc

# sin(y) = y * (1 + z*Q(z)), z = y^2 in [0, (pi/2)^2]
def fsin(z):
    if z == 0: return mp.mpf(-1)/6
    y = mp.sqrt(z); return (mp.sin(y)/y - 1)/z
for deg in (8, 9, 10):
    c = cheb_fit(fsin, mp.mpf(0), (mp.pi/2)**2 * mp.mpf('1.0001'), deg)
    ys = np.linspace(1e-6, np.pi/2, 200001)
    z = ys*ys
    approx = ys*(1 + z*horner_fp64(c, z))
    exact = np.array([float(mp.sin(mp.mpf(float(y)))) for y in ys[::50]])
    rel = np.abs(approx[::50]-exact)/exact
    print("sin deg", deg, "max rel err", rel.max()/2**-53, "ulp")
    if deg == 9: csin = c
# asin(x) = x*(1 + t*R(t)), t = x^2 in [0, 0.25]
def fasin(t):
    if t == 0: return mp.mpf(1)/6
    x = mp.sqrt(t); return (mp.asin(x)/x - 1)/t
for deg in (10, 11, 12, 13):
    c = cheb_fit(fasin, mp.mpf(0), mp.mpf('0.2501'), deg)
    xs = np.linspace(1e-8, 0.5, 100001)
    t = xs*xs
    approx = xs*(1 + t*horner_fp64(c, t))
    exact = np.array([float(mp.asin(mp.mpf(float(x)))) for x in xs[::25]])
    rel = np.abs(approx[::25]-exact)/exact
    print("asin deg", deg, "max rel err", rel.max()/2**-53, "ulp")
    if deg == 12: casin = c
# exp(r), |r| <= ln2/2 * 1.0001: e^r = 1 + r + r^2 * E(r)
def fexp(r):
    if r == 0: return mp.mpf(1)/2
    return (mp.exp(r) - 1 - r)/(r*r)
for deg in (9, 10, 11):
    L = mp.log(2)/2*mp.mpf('1.001')
    c = cheb_fit(fexp, -L, L, deg)
    rs = np.linspace(-float(L), float(L), 100001)
    approx = 1 + rs + rs*rs*horner_fp64(c, rs)
    exact = np.array([float(mp.exp(mp.mpf(float(r)))) for r in rs[::25]])
    rel = np.abs(approx[::25]-exact)/exact
    print("exp deg", deg, "max rel err", rel.max()/2**-53, "ulp")
    if deg == 10: cexp = c
def show(name, c):
    print(name, "= {", ", ".join(float(x).hex() for x in c), "};")
    print("  //", ", ".join(repr(float(x)) for x in c))
show("SIN_Q", csin); show("ASIN_R", casin); show("EXP_E", cexp)

print("---- pure approximation errors (mp evaluation) ----")
def approx_err(f_exact, build, lo, hi, n=4001):
    worst = mp.mpf(0)
    for i in range(1, n):
        x = lo + (hi-lo)*mp.mpf(i)/n
        e = abs(build(x) - f_exact(x))/abs(f_exact(x))
        worst = max(worst, e)
    return worst
def fexp2(r):
    if abs(r) < mp.mpf('1e-20'): return mp.mpf(1)/2 + r/6
    return (mp.exp(r) - 1 - r)/(r*r)
for deg in (6, 7, 8):
    c = cheb_fit(fsin, mp.mpf(0), (mp.pi/2)**2 * mp.mpf('1.0001'), deg)
    e = approx_err(mp.sin, lambda y: y*(1 + y*y*mp.polyval(c[::-1], y*y)), mp.mpf(0), mp.pi/2)
    print("sin Q deg", deg, mp.nstr(e/mp.mpf(2)**-53, 5), "ulp")
for deg in (9, 10, 11, 12):
    c = cheb_fit(fasin, mp.mpf(0), mp.mpf('0.2501'), deg)
    e = approx_err(mp.asin, lambda x: x*(1 + x*x*mp.polyval(c[::-1], x*x)), mp.mpf(0), mp.mpf('0.5'))
    print("asin R deg", deg, mp.nstr(e/mp.mpf(2)**-53, 5), "ulp")
for deg in (7, 8, 9):
    L = mp.log(2)/2*mp.mpf('1.001')
    c = cheb_fit(fexp2, -L, L, deg)
    e = approx_err(mp.exp, lambda r: 1 + r + r*r*mp.polyval(c[::-1], r), -L, L)
    print("exp E deg", deg, mp.nstr(e/mp.mpf(2)**-53, 5), "ulp")
    if deg == 9: show("EXP_E9", c)
    if deg == 8: show("EXP_E8", c)
c = cheb_fit(fsin, mp.mpf(0), (mp.pi/2)**2 * mp.mpf('1.0001'), 8); show("SIN_Q8", c)
c = cheb_fit(fsin, mp.mpf(0), (mp.pi/2)**2 * mp.mpf('1.0001'), 7); show("SIN_Q7", c)
c = cheb_fit(fasin, mp.mpf(0), mp.mpf('0.2501'), 11); show("ASIN_R11", c)
c = cheb_fit(fasin, mp.mpf(0), mp.mpf('0.2501'), 12); show("ASIN_R12", c)
