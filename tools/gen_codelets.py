#!/usr/bin/env python3
"""Generates cu-sdr-collection_b200/csrc/fft_codelets.cuh: straight-line register DFT codelets.

Every codelet works on a register array ``float2 (&x)[N]`` and hands each output to a functor
``emit(k, re, im)`` with a literal output index ``k`` (so stores/accumulations are resolved at
compile time after inlining).  All twiddles are literal constants, which lets ptxas use the
immediate form of FFMA.

  * odd primes P: direct DFT using the (x_j + x_{P-j}, x_j - x_{P-j}) symmetry — 4*((P-1)/2)^2 FMAs
  * powers of two: radix-2 decimation-in-frequency network, trivial twiddles special-cased,
    outputs emitted at their bit-reversed positions
  * composites of coprime factors (33 = 3 x 11): Good-Thomas prime-factor mapping inside the
    register file, so no twiddles between the two factors

Run:  python tools/gen_codelets.py > cu-sdr-collection_b200/csrc/fft_codelets.cuh
"""
import math
import sys

OUT = []


def w(s=""):
    OUT.append(s)


def lit(v):
    if abs(v) < 1e-17:
        return "0.0f"
    return repr(float(format(v, ".9g"))) + "f"


def prime_block(P, idx, omap, inv, tag, inplace=False):
    """DFT of odd prime length P over slots x[idx[j]]; output k -> emit(omap[k],..) or back in place."""
    h = (P - 1) // 2
    X = [f"x[{i}]" for i in idx]
    w(f"    {{ // DFT-{P} ({'inv' if inv else 'fwd'}) on slots {idx}")
    for j in range(1, h + 1):
        a, b = X[j], X[P - j]
        w(f"        {{ const float2 t = {a}; {a}.x = t.x + {b}.x; {a}.y = t.y + {b}.y; "
          f"{b}.x = t.x - {b}.x; {b}.y = t.y - {b}.y; }}")
    # X0
    sr = " + ".join([f"{X[0]}.x"] + [f"{X[j]}.x" for j in range(1, h + 1)])
    si = " + ".join([f"{X[0]}.y"] + [f"{X[j]}.y" for j in range(1, h + 1)])
    outs = {}
    w(f"        const float {tag}r0 = {sr};")
    w(f"        const float {tag}i0 = {si};")
    outs[0] = (f"{tag}r0", f"{tag}i0")
    for k in range(1, h + 1):
        ar, ai, br, bi = f"{tag}ar{k}", f"{tag}ai{k}", f"{tag}br{k}", f"{tag}bi{k}"
        w(f"        float {ar} = {X[0]}.x, {ai} = {X[0]}.y, {br}, {bi};")
        for j in range(1, h + 1):
            q = (j * k) % P
            c = math.cos(2 * math.pi * q / P)
            s = math.sin(2 * math.pi * q / P)
            w(f"        {ar} = fmaf({X[j]}.x, {lit(c)}, {ar}); {ai} = fmaf({X[j]}.y, {lit(c)}, {ai});")
            if j == 1:
                w(f"        {br} = {X[P - j]}.x * {lit(s)}; {bi} = {X[P - j]}.y * {lit(s)};")
            else:
                w(f"        {br} = fmaf({X[P - j]}.x, {lit(s)}, {br}); {bi} = fmaf({X[P - j]}.y, {lit(s)}, {bi});")
        # forward: X_k = A - iB ; X_{P-k} = A + iB.  inverse: swapped.
        lo = (f"{ar} + {bi}", f"{ai} - {br}")
        hi = (f"{ar} - {bi}", f"{ai} + {br}")
        if inv:
            lo, hi = hi, lo
        outs[k] = lo
        outs[P - k] = hi
        if not inplace:
            w(f"        emit({omap[k]}, {lo[0]}, {lo[1]});")
            w(f"        emit({omap[P - k]}, {hi[0]}, {hi[1]});")
    if not inplace:
        w(f"        emit({omap[0]}, {tag}r0, {tag}i0);")
    else:
        for k in range(P):
            w(f"        const float {tag}yr{k} = {outs[k][0]}, {tag}yi{k} = {outs[k][1]};")
        for k in range(P):
            w(f"        {X[k]}.x = {tag}yr{k}; {X[k]}.y = {tag}yi{k};")
    w("    }")


def pow2_block(N, inv):
    """radix-2 DIF network in place over x[0..N-1]; x[i] ends holding X[bitrev(i)]."""
    sgn = 1.0 if inv else -1.0
    span = N // 2
    while span >= 1:
        w(f"    // span {span}")
        for g in range(0, N, 2 * span):
            for j in range(span):
                a, b = f"x[{g + j}]", f"x[{g + j + span}]"
                ang = sgn * 2 * math.pi * j / (2 * span)
                wr, wi = math.cos(ang), math.sin(ang)
                w(f"    {{ const float tr = {a}.x - {b}.x, ti = {a}.y - {b}.y; {a}.x += {b}.x; {a}.y += {b}.y;")
                if j == 0:
                    w(f"      {b}.x = tr; {b}.y = ti; }}")
                elif 4 * j == 2 * span:      # w = -i (fwd) / +i (inv)
                    if inv:
                        w(f"      {b}.x = -ti; {b}.y = tr; }}")
                    else:
                        w(f"      {b}.x = ti; {b}.y = -tr; }}")
                elif 8 * j == 2 * span or 8 * j == 3 * 2 * span:
                    # w = (±1 ± i)/sqrt2
                    r2 = lit(math.sqrt(0.5))
                    sr = "+" if wr > 0 else "-"
                    # (tr + i ti)(wr + i wi) with |wr|=|wi|=r2
                    # real = tr*wr - ti*wi ; imag = tr*wi + ti*wr
                    re = f"({'' if wr > 0 else '-'}tr {'-' if wi > 0 else '+'} ti) * {r2}"
                    im = f"({'' if wi > 0 else '-'}tr {'+' if wr > 0 else '-'} ti) * {r2}"
                    w(f"      {b}.x = {re}; {b}.y = {im}; }}")
                else:
                    w(f"      {b}.x = fmaf(tr, {lit(wr)}, -ti * {lit(wi)}); {b}.y = fmaf(tr, {lit(wi)}, ti * {lit(wr)}); }}")
        span //= 2


def bitrev(i, n):
    r = 0
    b = n.bit_length() - 1
    for _ in range(b):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def gen_prime(P):
    for inv in (False, True):
        nm = f"dft{P}_{'inv' if inv else 'fwd'}"
        w(f"template <class F> __device__ __forceinline__ void {nm}(float2 (&x)[{P}], F&& emit)")
        w("{")
        prime_block(P, list(range(P)), list(range(P)), inv, "p")
        w("}")
        w()


def gen_pow2(N):
    for inv in (False, True):
        nm = f"dft{N}_{'inv' if inv else 'fwd'}"
        w(f"template <class F> __device__ __forceinline__ void {nm}(float2 (&x)[{N}], F&& emit)")
        w("{")
        pow2_block(N, inv)
        for i in range(N):
            w(f"    emit({bitrev(i, N)}, x[{i}].x, x[{i}].y);")
        w("}")
        w()


def egcd_inv(a, m):
    return pow(a, -1, m)


def gen_pfa(N1, N2):
    """N = N1*N2 coprime, N1 small prime done in place, N2 prime emitted (Good-Thomas)."""
    N = N1 * N2
    for inv in (False, True):
        nm = f"dft{N}_{'inv' if inv else 'fwd'}"
        w(f"// {N} = {N1} x {N2} prime-factor algorithm: input n = ({N2}*n1 + {N1}*n2) mod {N},")
        k1c = N2 * egcd_inv(N2, N1) % N
        k2c = N1 * egcd_inv(N1, N2) % N
        w(f"// output k = ({k1c}*k1 + {k2c}*k2) mod {N}; no twiddles between the two stages.")
        w(f"template <class F> __device__ __forceinline__ void {nm}(float2 (&x)[{N}], F&& emit)")
        w("{")
        for n2 in range(N2):
            idx = [(N2 * n1 + N1 * n2) % N for n1 in range(N1)]
            prime_block(N1, idx, None, inv, f"a{n2}_", inplace=True)
        for k1 in range(N1):
            idx = [(N2 * k1 + N1 * n2) % N for n2 in range(N2)]
            omap = [(k1c * k1 + k2c * k2) % N for k2 in range(N2)]
            prime_block(N2, idx, omap, inv, f"b{k1}_")
        w("}")
        w()


def main():
    w("// GENERATED by tools/gen_codelets.py — do not edit.  Register DFT codelets (fp32).")
    w("#pragma once")
    w("#include <cuda_runtime.h>")
    w()
    w("namespace gc { namespace codelet {")
    w()
    for P in (3, 5, 7, 11, 13, 31):
        gen_prime(P)
    for N in (2, 4, 8, 16, 32):
        gen_pow2(N)
    gen_pfa(3, 11)
    w("}}  // namespace gc::codelet")
    sys.stdout.write("\n".join(OUT) + "\n")


if __name__ == "__main__":
    main()
