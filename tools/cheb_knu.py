import numpy as np, mpmath as mp
mp.mp.dps = 40
def g(nu, x):  # sqrt(x) e^x K_nu(x)
    x = mp.mpf(x)
    return mp.sqrt(x) * mp.exp(x) * mp.besselk(nu, x)
def cheb_coeffs(f, a, b, n):
    k = np.arange(n)
    nodes = np.cos(np.pi * (k + 0.5) / n)            # in [-1, 1]
    t = 0.5 * (a + b) + 0.5 * (b - a) * nodes
    y = np.array([float(f(tt)) for tt in t])
    c = np.array([2.0 / n * np.sum(y * np.cos(np.pi * j * (k + 0.5) / n)) for j in range(n)])
    c[0] *= 0.5
    return c
def clenshaw(c, a, b, t):
    u = (2 * t - a - b) / (b - a)
    b1 = np.zeros_like(u); b2 = np.zeros_like(u)
    for cj in c[:0:-1]:
        b1, b2 = 2 * u * b1 - b2 + cj, b1
    return u * b1 - b2 + c[0]
segs = [(0.0, 0.25), (0.25, 0.5), (0.5, 1.0)]
for nu in (0.2, 0.75, 1.0, 1.25, 2.3, 3.2, 3.49):
    row = []
    for (a, b) in segs:
        f = lambda t: g(nu, 2.0 / t) if t > 0 else mp.sqrt(mp.pi / 2)
        best = None
        for n in (10, 12, 14, 16, 18, 20):
            c = cheb_coeffs(f, a, b, n)
            tt = np.linspace(a + 1e-9, b, 400)
            ex = np.array([float(f(v)) for v in tt])
            err = np.max(np.abs(clenshaw(c, a, b, tt) - ex) / ex)
            if err < 3e-15:
                best = (n, err); break
        row.append(best if best else ("none", err))
    print(nu, row)
