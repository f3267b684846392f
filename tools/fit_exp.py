import mpmath as mp
mp.mp.dps = 60
def fit(deg_q, half):
    # q(r) ~ (exp(r)-1-r)/r^2 on [-half, half], Chebyshev interpolation at deg_q+1 nodes
    n = deg_q + 1
    nodes = [half*mp.cos(mp.pi*(2*i+1)/(2*n)) for i in range(n)]
    f = lambda r: (mp.e**r - 1 - r)/r**2 if abs(r) > mp.mpf(10)**-20 else mp.mpf(1)/2 + r/6
    A = mp.matrix(n, n); b = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n): A[i, j] = x**j
        b[i] = f(x)
    c = mp.lu_solve(A, b)
    return [c[j] for j in range(n)]
def check(cs, half):
    # coefficients rounded to double; evaluate in high precision (rounding of arithmetic excluded)
    cd = [mp.mpf(float(c)) for c in cs]
    worst = 0
    for i in range(4001):
        r = -half + 2*half*i/4000
        q = sum(cd[j]*r**j for j in range(len(cd)))
        p = 1 + r + r*r*q
        e = abs(p/mp.e**r - 1)
        worst = max(worst, e)
    return worst
half = mp.log(2)/2
for dq in (8, 9, 10):
    cs = fit(dq, half)
    print("deg total", dq+2, "max rel err", mp.nstr(check(cs, half), 5))
    if dq == 9:
        for j, c in enumerate(cs): print("  c%d = %s" % (j+2, float(c).hex()), repr(float(c)))
half = mp.log(2)/64
for dq in (3, 4, 5):
    cs = fit(dq, half)
    print("J=5 table: deg total", dq+2, "max rel err", mp.nstr(check(cs, half), 5))
print("---- table variant deg 6 coefficients (q deg 4)")
half = mp.log(2)/64
cs = fit(4, half)
for j, c in enumerate(cs): print("  c%d = %s  // %r" % (j+2, float(c).hex(), float(c)))
print("ln2 =", float(mp.log(2)).hex(), " log2e =", float(1/mp.log(2)).hex(), " 32log2e=", float(32/mp.log(2)).hex(), " ln2/32=", float(mp.log(2)/32).hex())
half = mp.log(2)/2
cs = fit(9, half)
print("---- poly11")
for j, c in enumerate(cs): print("  c%d = %s  // %r" % (j+2, float(c).hex(), float(c)))
