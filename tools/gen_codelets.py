#!/usr/bin/env python3
"""Generates cu-sdr-collection_b200/csrc/fft_codelets.cuh: straight-line register DFT codelets
written in Blackwell's packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2).

A complex value is one float2 = one 64-bit register pair, so a complex add is ONE FADD2 and
"real constant times complex, accumulate" is ONE FFMA2 with a literal broadcast scalar.  The packed
instructions carry operand swizzles (`.LO_HI`, per-half negate), so a + i*b, a - i*b and a - b are
single instructions as well, and a multiplication by a literal complex twiddle is two
(FMUL2 by the real part, FFMA2 of the rotated value by the imaginary part).  The FMA pipe does
the same number of lane operations as the scalar form, but every floating-point instruction
takes one issue slot instead of two - and the issue slot is what bounds the transform kernels
(profiles/: 71 % issue-active against 62 % FMA pipe for the scalar codelets).

Every codelet works on a register array ``float2 (&x)[N]`` and hands each output to a functor
``emit(k, re, im)`` with a literal output index ``k``.  The code is emitted in SSA form (every
intermediate is a fresh ``const float2``); all index maps are resolved at generation time.

  * odd primes P: direct DFT with the (x_j + x_{P-j}, x_j - x_{P-j}) symmetry:
    A_k = x_0 + sum_j cos(2 pi jk/P) s_j,  B_k = sum_j sin(2 pi jk/P) d_j,  X_k = A_k -/+ i B_k
  * primes where it is cheaper (31: 417 packed operations against 510): Rader's cyclic convolution of length P - 1 through
    two (P-1)-point codelets and literal spectrum values
  * a multiplication by a literal complex constant (Cooley-Tukey twiddles, Rader's spectrum values) stays "lazy" until its
    consumer is known: where the product feeds a sum / difference pair it is accumulated straight into the sum and the
    difference is 2a - s (five operations instead of six for two products, three instead of four for one)
  * powers of two: radix-2 decimation in time (a +/- w*b as three FFMA2: u = b*(1 + i*t), a +/- m*u; w = 1, -i free)
  * composites: coprime factors by the Good-Thomas prime-factor mapping (no twiddles), repeated
    factors (9 = 3x3, 25 = 5x5) by Cooley-Tukey with literal twiddles.

Run:  python tools/gen_codelets.py > cu-sdr-collection_b200/csrc/fft_codelets.cuh
"""
import math
import sys

OUT = []
_n = [0]


def w(s=""):
    OUT.append(s)


def lit(v):
    if abs(v) < 1e-17:
        return "0.0f"
    return repr(float(format(v, ".9g"))) + "f"


def new(expr):
    _n[0] += 1
    name = f"v{_n[0]}"
    w(f"    const float2 {name} = {expr};")
    return name


def is_pow2(n):
    return n & (n - 1) == 0


def is_prime(n):
    return n > 1 and all(n % p for p in range(2, int(n ** 0.5) + 1))


def split(n):
    """n = n1 * n2: coprime split if one exists (prime-power part first), else p * (n/p)."""
    f = []
    m, p = n, 2
    while m > 1:
        if m % p == 0:
            q = 1
            while m % p == 0:
                q *= p
                m //= p
            f.append(q)
        p += 1
    if len(f) > 1:
        return f[0], n // f[0], True
    p = 2
    while n % p:
        p += 1
    return p, n // p, False


class Lazy:
    """v * (wr + i*wi) not yet emitted: a product that feeds a sum / difference pair is fused into it (pair_sd)."""

    def __init__(self, v, wr, wi):
        self.v, self.wr, self.wi, self.name = v, wr, wi, None


def force(x):
    if isinstance(x, Lazy):
        if x.name is None:
            x.name = new(f"f2cmulc({x.v}, {lit(x.wr)}, {lit(x.wi)})")
        return x.name
    return x


def cmul_const(v, wr, wi):
    """v * (wr + i*wi) with literal wr, wi (general constants stay lazy until their consumer is known)."""
    v = force(v)
    if abs(wi) < 1e-15:
        if abs(wr - 1) < 1e-15:
            return v
        return new(f"f2scale({v}, {lit(wr)})")
    if abs(wr) < 1e-15:
        return new(f"f2scale(f2rot({v}), {lit(wi)})")
    return Lazy(v, wr, wi)


def pair_sd(a, b):
    """(a + b, a - b).  A literal product among the operands is accumulated straight into the sum (two FFMA2) and the difference
    taken as 2a - s (or s - 2b): five operations instead of six for two products, three instead of four for one."""
    la, lb = isinstance(a, Lazy) and a.name is None, isinstance(b, Lazy) and b.name is None
    if lb:
        pa = force(a)
        t = new(f"f2fma({b.v}, {lit(b.wr)}, {pa})")
        sm = new(f"f2fma(f2rot({b.v}), {lit(b.wi)}, {t})")
        return sm, new(f"f2fma({pa}, 2.0f, f2neg({sm}))")
    if la:
        pb = force(b)
        t = new(f"f2fma({a.v}, {lit(a.wr)}, {pb})")
        sm = new(f"f2fma(f2rot({a.v}), {lit(a.wi)}, {t})")
        return sm, new(f"f2fma({pb}, -2.0f, {sm})")
    a, b = force(a), force(b)
    return new(f"f2add({a}, {b})"), new(f"f2sub({a}, {b})")


def prime_block(P, vals, inv):
    h = (P - 1) // 2
    vals = list(vals)
    vals[0] = force(vals[0])
    s, d = [None], [None]
    for j in range(1, h + 1):
        sj, dj = pair_sd(vals[j], vals[P - j])
        s.append(sj)
        d.append(dj)
    out = [None] * P
    acc = vals[0]
    for j in range(1, h + 1):
        acc = new(f"f2add({acc}, {s[j]})")
    out[0] = acc
    for k in range(1, h + 1):
        # A + iB as ONE chain (the sine terms enter as i*d_j through the operand swizzle), A - iB = 2A - (A + iB):
        # 2h + 1 packed operations per output pair instead of 2h + 2
        a = vals[0]
        for j in range(1, h + 1):
            q = (j * k) % P
            a = new(f"f2fma({s[j]}, {lit(math.cos(2 * math.pi * q / P))}, {a})")
        hi = a
        for j in range(1, h + 1):
            q = (j * k) % P
            hi = new(f"f2fma(f2rot({d[j]}), {lit(math.sin(2 * math.pi * q / P))}, {hi})")
        lo = new(f"f2fma({a}, 2.0f, f2neg({hi}))")        # A - iB: forward X_k, inverse X_{P-k}
        if inv:
            lo, hi = hi, lo
        out[k] = lo
        out[P - k] = hi
    return out


def packed_ops_direct(P):
    h = (P - 1) // 2
    return 2 * h + h + h * (2 * h + 1)


def packed_ops(N):
    """packed operations of the codelet dft(N) emits (used to choose between the direct and Rader forms of a prime)."""
    if N == 1:
        return 0
    if is_pow2(N):
        bits = N.bit_length() - 1
        total, half = 0, 1
        while half < N:
            triv = 2 if half >= 2 else 1
            total += (N // (2 * half)) * (min(triv, half) * 2 + max(0, half - triv) * 3)
            half *= 2
        return total
    if is_prime(N):
        return min(packed_ops_direct(N), packed_ops_rader(N))
    n1, n2, coprime = split(N)
    return n2 * packed_ops(n1) + n1 * packed_ops(n2) + (0 if coprime else 2 * (n1 - 1) * (n2 - 1))


def packed_ops_rader(P):
    return 2 * packed_ops(P - 1) + 2 * (P - 2) + 2


def primitive_root(P):
    for g in range(2, P):
        x, seen = 1, set()
        for _ in range(P - 1):
            x = x * g % P
            seen.add(x)
        if len(seen) == P - 1:
            return g
    raise ValueError(P)


def rader_block(P, vals, inv):
    """Rader: X_{g^-m} = x_0 + (a (*) b)_m with a_q = x_{g^q}, b_r = w^(g^-r), the cyclic convolution of length P - 1 through
    two (P-1)-point codelets and P - 1 literal spectrum values; x_0 enters through the DC term of the product."""
    import cmath
    n = P - 1
    g = primitive_root(P)
    gi = pow(g, -1, P)
    sgn = 1.0 if inv else -1.0
    b = [cmath.exp(sgn * 2j * math.pi * pow(gi, r, P) / P) for r in range(n)]
    Bf = [sum(b[r] * cmath.exp(-2j * math.pi * k * r / n) for r in range(n)) / n for k in range(n)]
    vals = [force(v) for v in vals]
    a = [vals[pow(g, q, P)] for q in range(n)]
    A = dft(n, a, False)
    out = [None] * P
    out[0] = new(f"f2add({vals[0]}, {A[0]})")
    Cs = [new(f"f2fma({A[0]}, {lit(Bf[0].real)}, {vals[0]})")]          # Bf[0] = -1/n (real): C_0 + x_0
    assert abs(Bf[0].imag) < 1e-12
    for k in range(1, n):
        Cs.append(cmul_const(A[k], Bf[k].real, Bf[k].imag))
    c = dft(n, Cs, True)
    for m in range(n):
        out[pow(gi, m, P)] = c[m]
    return out


def bitrev(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def pow2_block(N, vals, inv):
    """radix-2 decimation in time."""
    sgn = 1.0 if inv else -1.0
    bits = N.bit_length() - 1
    cur = [vals[bitrev(i, bits)] for i in range(N)]
    half = 1
    while half < N:
        nxt = list(cur)
        for g in range(0, N, 2 * half):
            for j in range(half):
                a, b = cur[g + j], cur[g + j + half]
                num, den = j, 2 * half                     # w = exp(sgn * 2 pi i * j / (2*half))
                if num == 0:
                    nxt[g + j], nxt[g + j + half] = pair_sd(a, b)
                    continue
                a, b = force(a), force(b)
                if 4 * num == den:                       # w = -i (fwd) / +i (inv)
                    if inv:
                        nxt[g + j] = new(f"f2addi({a}, {b})")
                        nxt[g + j + half] = new(f"f2subi({a}, {b})")
                    else:
                        nxt[g + j] = new(f"f2subi({a}, {b})")
                        nxt[g + j + half] = new(f"f2addi({a}, {b})")
                else:
                    # a +/- w*b in THREE packed operations instead of four: the twiddle is factored as
                    # w = wr * (1 + i*t) or w = wi * (r + i) with the ratio of the smaller to the larger part
                    # (|ratio| <= 1), u = b + i*t*b (or r*b + i*b) is one FFMA2 shared by both outputs
                    ang = sgn * 2 * math.pi * num / den
                    wr, wi = math.cos(ang), math.sin(ang)
                    if abs(wr) >= abs(wi):
                        u = new(f"f2fma(f2rot({b}), {lit(wi / wr)}, {b})")
                        m = wr
                    else:
                        u = new(f"f2fma({b}, {lit(wr / wi)}, f2rot({b}))")
                        m = wi
                    nxt[g + j] = new(f"f2fma({u}, {lit(m)}, {a})")
                    nxt[g + j + half] = new(f"f2fma({u}, {lit(-m)}, {a})")
        cur = nxt
        half *= 2
    return cur


def dft(N, vals, inv):
    """DFT-N of the SSA values vals[0..N-1]; returns the outputs in natural order."""
    if N == 1:
        return [force(v) for v in vals]
    if is_pow2(N):
        return pow2_block(N, vals, inv)
    if is_prime(N):
        if packed_ops_rader(N) < packed_ops_direct(N):
            return rader_block(N, vals, inv)
        return prime_block(N, vals, inv)
    n1, n2, coprime = split(N)
    sgn = 1.0 if inv else -1.0
    out = [None] * N
    if coprime:
        # Good-Thomas: n = (n2*a + n1*b) mod N ; k = (c1*k1 + c2*k2) mod N
        c1 = n2 * pow(n2, -1, n1) % N
        c2 = n1 * pow(n1, -1, n2) % N
        cols = [dft(n1, [vals[(n2 * a + n1 * b) % N] for a in range(n1)], inv) for b in range(n2)]
        for k1 in range(n1):
            row = dft(n2, [cols[b][k1] for b in range(n2)], inv)
            for k2 in range(n2):
                out[(c1 * k1 + c2 * k2) % N] = row[k2]
        return out
    # Cooley-Tukey with literal twiddles: n = n2*a + b ; k = k1 + n1*k2
    cols = [dft(n1, [vals[n2 * a + b] for a in range(n1)], inv) for b in range(n2)]
    for k1 in range(n1):
        rowin = []
        for b in range(n2):
            ang = sgn * 2 * math.pi * k1 * b / N
            rowin.append(cmul_const(cols[b][k1], math.cos(ang), math.sin(ang)))
        row = dft(n2, rowin, inv)
        for k2 in range(n2):
            out[k1 + n1 * k2] = row[k2]
    return out


def gen_codelet(N):
    for inv in (False, True):
        nm = f"dft{N}_{'inv' if inv else 'fwd'}"
        w(f"template <class F> __device__ __forceinline__ void {nm}(float2 (&x)[{N}], F&& emit)")
        w("{")
        res = [force(r) for r in dft(N, [f"x[{i}]" for i in range(N)], inv)]
        for k in range(N):
            w(f"    emit({k}, {res[k]}.x, {res[k]}.y);")
        w("}")
        w()


PRELUDE = r"""
// packed fp32x2 primitives (one instruction each on sm_100a; the swizzled / negated operands fold
// into the FADD2 / FFMA2 operand modifiers)
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 f2addi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }   // a + i*b
__device__ __forceinline__ float2 f2subi(float2 a, float2 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }   // a - i*b
__device__ __forceinline__ float2 f2rot(float2 a) { return make_float2(-a.y, a.x); }                              // i*a
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 f2scale(float2 a, float c) { return __fmul2_rn(a, make_float2(c, c)); }
__device__ __forceinline__ float2 f2fma(float2 a, float c, float2 b) { return __ffma2_rn(a, make_float2(c, c), b); }   // a*c + b
// t * (wr + i*wi)
__device__ __forceinline__ float2 f2cmulc(float2 t, float wr, float wi)
{
    return __ffma2_rn(f2rot(t), make_float2(wi, wi), __fmul2_rn(t, make_float2(wr, wr)));
}
// a + b * (wr + i*wi)
__device__ __forceinline__ float2 f2bfly(float2 a, float2 b, float wr, float wi)
{
    return __ffma2_rn(f2rot(b), make_float2(wi, wi), __ffma2_rn(b, make_float2(wr, wr), a));
}
"""


def main():
    w("// GENERATED by tools/gen_codelets.py - do not edit.  Register DFT codelets in packed fp32x2 arithmetic.")
    w("#pragma once")
    w("#ifndef CODELET_HOST_CHECK   // tests/test_host.py compiles this header for the host with stub intrinsics")
    w("#include <cuda_runtime.h>")
    w("#endif")
    w()
    w("namespace gc { namespace codelet {")
    w(PRELUDE)
    sizes = (4, 8, 9, 10, 16, 18, 20, 25, 30, 31, 32, 33, 40, 45, 50)
    for N in sizes:
        gen_codelet(N)
    w("// compile-time dispatch: dft<N, INV>(x, emit)")
    w("template <int N, bool INV, class F> __device__ __forceinline__ void dft(float2 (&x)[N], F&& emit)")
    w("{")
    for i, N in enumerate(sizes):
        w(f"    {'if' if i == 0 else 'else if'} constexpr (N == {N}) {{ if constexpr (INV) dft{N}_inv(x, emit); else dft{N}_fwd(x, emit); }}")
    w("    else static_assert(N < 0, \"no codelet for this length\");")
    w("}")
    w()
    w("}}  // namespace gc::codelet")
    sys.stdout.write("\n".join(OUT) + "\n")


if __name__ == "__main__":
    main()
