#!/usr/bin/env python3
"""Generates cu-sdr-collection_b200/csrc/fft_codelets.cuh: straight-line register DFT codelets.

Every codelet works on a register array ``float2 (&x)[N]`` and hands each output to a functor
``emit(k, re, im)`` with a literal output index ``k`` (so stores/accumulations are resolved at
compile time after inlining).  All twiddles are literal constants, which lets ptxas use the
immediate form of FFMA.

  * odd primes P: direct DFT using the (x_j + x_{P-j}, x_j - x_{P-j}) symmetry - 4*((P-1)/2)^2 FMAs
  * powers of two: radix-2 decimation-in-frequency network, trivial twiddles special-cased
  * composites: recursive - coprime factors by the Good-Thomas prime-factor mapping (no twiddles),
    repeated factors (9 = 3x3, 25 = 5x5) by Cooley-Tukey with literal twiddles; all index maps are
    resolved at generation time, so the emitted code is flat register arithmetic.

Run:  python tools/gen_codelets.py > cu-sdr-collection_b200/csrc/fft_codelets.cuh
"""
import math
import sys

OUT = []
_uid = [0]


def w(s=""):
    OUT.append(s)


def uid():
    _uid[0] += 1
    return f"t{_uid[0]}_"


def lit(v):
    if abs(v) < 1e-17:
        return "0.0f"
    return repr(float(format(v, ".9g"))) + "f"


def is_pow2(n):
    return n & (n - 1) == 0


def is_prime(n):
    return n > 1 and all(n % p for p in range(2, int(n ** 0.5) + 1))


def bitrev(i, n):
    r = 0
    for _ in range(n.bit_length() - 1):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


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


def prime_block(P, slots, inv, omap):
    """DFT of odd prime length P over x[slots[j]].  omap = list -> emit(omap[k]); None -> in place
    (returns pos with x[slots[pos[k]]] holding output k)."""
    tag = uid()
    h = (P - 1) // 2
    X = [f"x[{i}]" for i in slots]
    w(f"    {{ // DFT-{P} ({'inv' if inv else 'fwd'}) on slots {slots}")
    for j in range(1, h + 1):
        a, b = X[j], X[P - j]
        w(f"        {{ const float2 t = {a}; {a}.x = t.x + {b}.x; {a}.y = t.y + {b}.y; "
          f"{b}.x = t.x - {b}.x; {b}.y = t.y - {b}.y; }}")
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
        lo = (f"{ar} + {bi}", f"{ai} - {br}")      # forward: X_k = A - iB ; X_{P-k} = A + iB
        hi = (f"{ar} - {bi}", f"{ai} + {br}")
        if inv:
            lo, hi = hi, lo
        outs[k] = lo
        outs[P - k] = hi
        if omap is not None:
            w(f"        emit({omap[k]}, {lo[0]}, {lo[1]});")
            w(f"        emit({omap[P - k]}, {hi[0]}, {hi[1]});")
    if omap is not None:
        w(f"        emit({omap[0]}, {tag}r0, {tag}i0);")
        w("    }")
        return None
    for k in range(P):
        w(f"        const float {tag}yr{k} = {outs[k][0]}, {tag}yi{k} = {outs[k][1]};")
    for k in range(P):
        w(f"        {X[k]}.x = {tag}yr{k}; {X[k]}.y = {tag}yi{k};")
    w("    }")
    return list(range(P))


def pow2_block(N, slots, inv, omap):
    """radix-2 DIF network in place over x[slots[i]]; x[slots[i]] ends holding X[bitrev(i)]."""
    sgn = 1.0 if inv else -1.0
    X = [f"x[{i}]" for i in slots]
    span = N // 2
    w(f"    // DFT-{N} ({'inv' if inv else 'fwd'}) radix-2 DIF on slots {slots}")
    while span >= 1:
        for g in range(0, N, 2 * span):
            for j in range(span):
                a, b = X[g + j], X[g + j + span]
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
                    r2 = lit(math.sqrt(0.5))
                    re = f"({'' if wr > 0 else '-'}tr {'-' if wi > 0 else '+'} ti) * {r2}"
                    im = f"({'' if wi > 0 else '-'}tr {'+' if wr > 0 else '-'} ti) * {r2}"
                    w(f"      {b}.x = {re}; {b}.y = {im}; }}")
                else:
                    w(f"      {b}.x = fmaf(tr, {lit(wr)}, -ti * {lit(wi)}); {b}.y = fmaf(tr, {lit(wi)}, ti * {lit(wr)}); }}")
        span //= 2
    pos = [0] * N
    for i in range(N):
        pos[bitrev(i, N)] = i
    if omap is not None:
        for k in range(N):
            w(f"    emit({omap[k]}, {X[pos[k]]}.x, {X[pos[k]]}.y);")
        return None
    return pos


def gen(N, slots, inv, omap):
    """DFT-N over x[slots[n]], n = 0..N-1.  omap: emit outputs as emit(omap[k], ..); None: in place,
    returns pos (x[slots[pos[k]]] = output k)."""
    if N == 1:
        if omap is not None:
            w(f"    emit({omap[0]}, x[{slots[0]}].x, x[{slots[0]}].y);")
            return None
        return [0]
    if is_pow2(N):
        return pow2_block(N, slots, inv, omap)
    if is_prime(N):
        return prime_block(N, slots, inv, omap)
    n1, n2, coprime = split(N)
    sgn = 1.0 if inv else -1.0
    if coprime:
        # Good-Thomas: n = (n2*a + n1*b) mod N ; k = (c1*k1 + c2*k2) mod N
        c1 = n2 * pow(n2, -1, n1) % N
        c2 = n1 * pow(n1, -1, n2) % N
        pos1 = {}
        for b in range(n2):
            sub = [slots[(n2 * a + n1 * b) % N] for a in range(n1)]
            pos1[b] = gen(n1, sub, inv, None)
        out_pos = [None] * N
        for k1 in range(n1):
            sub = [slots[(n2 * pos1[b][k1] + n1 * b) % N] for b in range(n2)]
            if omap is not None:
                gen(n2, sub, inv, [omap[(c1 * k1 + c2 * k2) % N] for k2 in range(n2)])
            else:
                p2 = gen(n2, sub, inv, None)
                for k2 in range(n2):
                    b = p2[k2]
                    out_pos[(c1 * k1 + c2 * k2) % N] = (n2 * pos1[b][k1] + n1 * b) % N
        return None if omap is not None else out_pos
    # Cooley-Tukey with literal twiddles: n = n2*a + b ; k = k1 + n1*k2
    pos1 = {}
    for b in range(n2):
        sub = [slots[n2 * a + b] for a in range(n1)]
        pos1[b] = gen(n1, sub, inv, None)
    for b in range(1, n2):
        for k1 in range(1, n1):
            ang = sgn * 2 * math.pi * k1 * b / N
            wr, wi = math.cos(ang), math.sin(ang)
            s = f"x[{slots[n2 * pos1[b][k1] + b]}]"
            w(f"    {{ const float tr = {s}.x, ti = {s}.y; {s}.x = fmaf(tr, {lit(wr)}, -ti * {lit(wi)}); "
              f"{s}.y = fmaf(tr, {lit(wi)}, ti * {lit(wr)}); }}")
    out_pos = [None] * N
    for k1 in range(n1):
        sub = [slots[n2 * pos1[b][k1] + b] for b in range(n2)]
        if omap is not None:
            gen(n2, sub, inv, [omap[k1 + n1 * k2] for k2 in range(n2)])
        else:
            p2 = gen(n2, sub, inv, None)
            for k2 in range(n2):
                b = p2[k2]
                out_pos[k1 + n1 * k2] = n2 * pos1[b][k1] + b
    return None if omap is not None else out_pos


def gen_codelet(N):
    for inv in (False, True):
        nm = f"dft{N}_{'inv' if inv else 'fwd'}"
        w(f"template <class F> __device__ __forceinline__ void {nm}(float2 (&x)[{N}], F&& emit)")
        w("{")
        gen(N, list(range(N)), inv, list(range(N)))
        w("}")
        w()


def main():
    w("// GENERATED by tools/gen_codelets.py - do not edit.  Register DFT codelets (fp32).")
    w("#pragma once")
    w("#include <cuda_runtime.h>")
    w()
    w("namespace gc { namespace codelet {")
    w()
    sizes = (2, 3, 4, 5, 7, 8, 9, 11, 13, 16, 25, 30, 31, 32, 33, 40, 45, 50)
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
