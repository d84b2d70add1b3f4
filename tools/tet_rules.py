"""Constants of the two tetrahedron quadrature rules the PNP-Stokes forms need (the rules FFC picked for
benchmarks/pnp_stokes/vector_linear_pnp_ns_forms.h: degree 5 / 15 points at :16414, degree 4 / 14 points at :21776),
derived here instead of being copied from the generated header:
  * 15 points (Keast): closed forms — weights 6544/36015, 81/2240, 161051/2304960, 338/5145 (x 1/6), orbit
    parameters 1/4, 1/3, 1/11, (1 - sqrt(7/13))/4;
  * 14 points: the 6 edge midpoints (weight 1/315) and two vertex orbits (a, a, a, 1 - 3a) whose (a, w) solve the
    moment equations of degree 0, 2, 3, 4.
Prints 20 digits; csrc/ns_forms.cuh and oracle/ns_elements.cpp hold the values.   python tools/tet_rules.py"""
from mpmath import mp, mpf, findroot, factorial, sqrt

mp.dps = 40
w0 = mpf(1) / 315


def mom(k):          # integral of lambda_1^k over the reference tetrahedron (volume 1/6)
    return factorial(k) / factorial(k + 3)


def eqs(a, w1, b, w2):
    out = []
    for k in (0, 2, 3, 4):
        s = w0 * (3 * (mpf(1) / 2) ** k + (3 if k == 0 else 0))
        s += w1 * (3 * a ** k + (1 - 3 * a) ** k) + w2 * (3 * b ** k + (1 - 3 * b) ** k)
        out.append(s - mom(k))
    return out


a, w1, b, w2 = findroot(eqs, (mpf("0.1005"), mpf("0.01476"), mpf("0.31437"), mpf("0.02214")))
print("14-point rule: a1 %s w1 %s a2 %s w2 %s" % tuple(mp.nstr(v, 20) for v in (a, w1, b, w2)))
print("15-point rule: w %s %s %s %s  edge-orbit a %s" % tuple(mp.nstr(v, 20) for v in (
    mpf(6544) / 36015 / 6, mpf(81) / 2240 / 6, mpf(161051) / 2304960 / 6, mpf(338) / 5145 / 6, (1 - sqrt(mpf(7) / 13)) / 4)))
