"""Development tool: the quotient step q' = RN(q + r y), q = RN(a y), r = a - b q, y = RN(1 / b), against exact rational arithmetic
for the odd divisors b = 1 .. 63 of nmath's bd0 series (drj_div_odd, drj_device.cuh): 768,000 samples, 0 mismatches."""
from fractions import Fraction as F
import random, struct
random.seed(1)
def rnd_double():
    # random mantissa, exponent in a moderate range, random sign
    m = random.getrandbits(52) | (1 << 52)
    e = random.randint(-60, 60)
    x = float(m) * 2.0 ** (e - 52)
    return x if random.random() < 0.5 else -x
bad = 0; n = 0
for j in range(32):
    b = 2 * j + 1
    y = 1.0 / b
    for _ in range(15000):
        a = rnd_double()
        q = a * y                                  # RN(a*y)
        r = float(F(a) - F(b) * F(q))              # fma(-b, q, a)
        qp = float(F(q) + F(r) * F(y))             # fma(r, y, q)
        want = float(F(a) / b)
        n += 1
        if qp != want:
            bad += 1
            if bad < 10: print("MISMATCH", j, a.hex(), qp.hex(), want.hex())
    # adversarial: a = k * b +- tiny and mantissas near all-ones
    for _ in range(3000):
        k = random.getrandbits(53) | (1 << 52)
        a = float(k)
        for a2 in (a, a * b if a * b < 2**1000 else a, struct.unpack('<d', struct.pack('<Q', struct.unpack('<Q', struct.pack('<d', a))[0] | ((1 << 20) - 1)))[0]):
            q = a2 * y; r = float(F(a2) - F(b) * F(q)); qp = float(F(q) + F(r) * F(y)); want = float(F(a2) / b); n += 1
            if qp != want:
                bad += 1
                if bad < 10: print("MISMATCH adv", j, a2.hex(), qp.hex(), want.hex())
print("checked", n, "mismatches", bad)
