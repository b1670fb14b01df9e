#!/usr/bin/env python3
"""Regenerate Go's math/rand `rngCooked[607]` seeding table without a Go toolchain.

Go's stdlib (go1.24, src/math/rand/rng.go) seeds its additive lagged-Fibonacci
source by XOR-ing a Lehmer stream with a fixed 607-word table, `rngCooked`.  The
table is defined (src/math/rand/gen_cooked.go) as the generator state after
7.8e12 steps of the same ALFG recurrence, started from a Lehmer-seeded state
(seed 1, shifts 20/10, no table).  The recurrence is linear over Z/2^64:

    s_n = s_{n-607} + s_{n-273}            (mod 2^64)

so the state after N steps is obtained by computing x^n mod (x^607 - x^334 - 1)
with square-and-multiply instead of iterating 7.8e12 times.

Outputs `go_rng_cooked.inc` (607 int64 literals) for both the oracle and the
product host mapgen, and prints fingerprints that SURVEY.md Appendix B records
(first/last words, SHA-256, sum) plus canonical Go outputs for Seed(1).
"""
import hashlib
import os
import sys

import numpy as np

LEN, TAP = 607, 273
M31 = (1 << 31) - 1
MASK64 = (1 << 64) - 1
N_STEPS = 7_800_000_000_000  # 7.8e12, gen_cooked.go main()


def seedrand(x):
    # x[n+1] = 48271 * x[n] mod (2**31 - 1)   (rng.go seedrand)
    hi, lo = divmod(x, 44488)
    x = 48271 * lo - 3399 * hi
    if x < 0:
        x += M31
    return x


def lehmer_state(seed, sh_hi, sh_mid):
    seed %= M31
    if seed < 0:
        seed += M31
    if seed == 0:
        seed = 89482311
    x = seed
    vec = [0] * LEN
    for i in range(-20, LEN):
        x = seedrand(x)
        if i >= 0:
            u = (x << sh_hi) & MASK64
            x = seedrand(x)
            u ^= (x << sh_mid) & MASK64
            x = seedrand(x)
            u ^= x
            vec[i] = u
    return vec


def polymul(a, b):
    """(a*b) mod (x^607 - x^334 - 1) over Z/2^64; a, b uint64 arrays of 607."""
    full = np.zeros(2 * LEN - 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(LEN):
            ai = a[i]
            if ai:
                full[i:i + LEN] += ai * b
        # x^k = x^(k-273) + x^(k-607) for k >= 607, highest degree first
        for k in range(2 * LEN - 2, LEN - 1, -1):
            c = full[k]
            if c:
                full[k - TAP] += c
                full[k - LEN] += c
                full[k] = 0
    return full[:LEN].copy()


def mulx(a):
    """a*x mod P."""
    out = np.zeros(LEN, dtype=np.uint64)
    out[1:] = a[:-1]
    top = a[-1]
    with np.errstate(over="ignore"):
        out[LEN - TAP] += top  # x^607 -> x^334 + 1
        out[0] += top
    return out


def xpow(n):
    result = np.zeros(LEN, dtype=np.uint64)
    result[0] = 1
    base = np.zeros(LEN, dtype=np.uint64)
    base[1] = 1
    while n:
        if n & 1:
            result = polymul(result, base)
        n >>= 1
        if n:
            base = polymul(base, base)
    return result


def cooked_table():
    vec0 = lehmer_state(1, 20, 10)
    # s_n for n = -606..0 sit at index (334 - n) mod 607
    base = np.zeros(LEN, dtype=np.uint64)
    for k in range(LEN):
        n = k - 606
        base[k] = vec0[(334 - n) % LEN]
    out = [0] * LEN
    m0 = N_STEPS - 606
    poly = xpow(m0 + 606)
    with np.errstate(over="ignore"):
        for m in range(m0, N_STEPS + 1):
            s_m = int(np.sum(poly * base, dtype=np.uint64))
            out[(334 - m) % LEN] = s_m
            poly = mulx(poly)
    return out


class GoRand:
    def __init__(self, cooked, seed):
        self.vec = [0] * LEN
        self.tap, self.feed = 0, LEN - TAP
        seed %= M31
        if seed < 0:
            seed += M31
        if seed == 0:
            seed = 89482311
        x = seed
        for i in range(-20, LEN):
            x = seedrand(x)
            if i >= 0:
                u = (x << 40) & MASK64
                x = seedrand(x)
                u ^= (x << 20) & MASK64
                x = seedrand(x)
                u ^= x
                self.vec[i] = u ^ cooked[i]

    def uint64(self):
        self.tap = (self.tap - 1) % LEN
        self.feed = (self.feed - 1) % LEN
        x = (self.vec[self.feed] + self.vec[self.tap]) & MASK64
        self.vec[self.feed] = x
        return x

    def int63(self):
        return self.uint64() & ((1 << 63) - 1)

    def int31(self):
        return self.int63() >> 32

    def intn(self, n):
        if n & (n - 1) == 0:
            return self.int31() & (n - 1)
        mx = (1 << 31) - 1 - (1 << 31) % n
        v = self.int31()
        while v > mx:
            v = self.int31()
        return v % n


def main():
    cooked = cooked_table()
    signed = [c - (1 << 64) if c >= (1 << 63) else c for c in cooked]
    blob = b"".join(c.to_bytes(8, "little") for c in cooked)
    sha = hashlib.sha256(blob).hexdigest()
    print("rngCooked[0..3] =", signed[:4])
    print("rngCooked[-2:]  =", signed[-2:])
    print("sha256(le u64)  =", sha)
    print("sum mod 2^64    =", sum(cooked) & MASK64)
    r = GoRand(cooked, 1)
    print("Seed(1) Int63   =", [r.int63() for _ in range(3)])
    r = GoRand(cooked, 1)
    print("Seed(1) Intn(100) x10 =", [r.intn(100) for _ in range(10)])
    r = GoRand(cooked, 12345)
    print("Seed(12345) Intn(20) x8 =", [r.intn(20) for _ in range(8)])

    expect_sha = "1928503b93a563e491119a5889baba73d1605b90b63634c14a408805797a7c7b"
    r = GoRand(cooked, 1)
    canonical = [r.intn(100) for _ in range(10)] == [81, 87, 47, 59, 81, 18, 25, 40, 56, 0]
    if sha != expect_sha or not canonical:
        print("FINGERPRINT MISMATCH", file=sys.stderr)
        sys.exit(1)

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = (
        "/* Go math/rand rngCooked[607] (go1.24 src/math/rand/rng.go), regenerated by\n"
        " * tools/gen_rngcooked.py via ALFG jump-ahead; sha256(le u64) =\n"
        " * %s */\n" % sha
    )
    body = "".join("  %dLL,\n" % s if s != -(1 << 63) else "  (-9223372036854775807LL-1),\n" for s in signed)
    for rel in ("oracle/go_rng_cooked.inc", "generalsreinforcementlearning_b200/csrc/go_rng_cooked.inc"):
        with open(os.path.join(root, rel), "w") as f:
            f.write(header + body)
        print("wrote", rel)


if __name__ == "__main__":
    main()


def lehmer_pow_table(path):
    """48271^k mod (2^31-1), k = 0..1841: the device map generator seeds Go's generator with independent
    multiplications x_k = 48271^k * x_0 instead of 1,841 dependent Lehmer steps (grl_mapgen_gpu.cu)."""
    p, a, vals = 2 ** 31 - 1, 48271, [1]
    for _ in range(1841):
        vals.append(vals[-1] * a % p)
    with open(path, "w") as f:
        f.write("// 48271^k mod (2^31 - 1) for k = 0..1841: jump-ahead table of the Lehmer generator Go's math/rand seeds with\n"
                "// (rng.go seedrand); generated by tools/gen_rngcooked.py --lehmer-pow\n")
        for i in range(0, len(vals), 8):
            f.write(", ".join(str(v) for v in vals[i:i + 8]) + ",\n")


if __name__ == "__main__" and "--lehmer-pow" in __import__("sys").argv:
    import os
    lehmer_pow_table(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "generalsreinforcementlearning_b200", "csrc",
                                  "go_rng_lehmer_pow.inc"))
