// Primary-code generators of the signals whose codes the reference builds at run time, as bit-packed shift
// registers.  Every function is __host__ __device__: the library runs them in codegen_kernel (one thread per SV and component,
// gc_generate_code_device and the engine's own code set-up) and on the host (gc_generate_code).  The ICD constant tables are data
// extracted from the reference tree (icd_tables.inc, tools/extract_icd_tables.py).  Outputs are +-1 chips (logic 0 -> +1, 1 -> -1)
// in the layout gc_set_code takes for the signal.
//
//   GPS L5 I5 / Q5      GPS/GPS_L5C/include/generateL5Icode.m:44-133, generateL5Qcode.m   (XA short-cycled at 8190, XB advanced)
//   GAL E5a / E5b I, Q  GAL/GAL_E5a/include/generateE5aIcode.m:45-108 and its three twins  (two 14-stage registers, octal taps)
//   GAL E5a-Q / E5b-Q secondary codes   generateE5aQ_secondary.m:73-87                     (25 hex characters -> 100 chips)
//   BDS B2a data / pilot  BDS/B2a/include/generateB2aDataCode.m:111-138, generateB2aPilotCode.m  (register 1 reset after 8190 chips)
//   BDS B1I             BDS/B1I/include/generateCAcode53.m:38-103                           (G1, G2 with 2 or 3 phase-selector taps)
//   GPS L2C CM / CL     GPS/GPS_L2C/include/generateCMcode.m:88-111, generateCLcode.m       (27-stage modular register, return to zero)
//   BDS B1C data / pilot BOC(1,1), pilot BOC(6,1)   BDS/B1C/include/generateDataBOC11.m:66-90, generatePilotBOC11.m,
//                       generatePilotBOC61.m:103-110, JacobiSymbol.m                        (Weil codes from the Legendre sequence of 10243)
//   GAL E1-B / E1-C     GAL/GAL_E1C/include/generateE1Bcode.m:44-55 with include/E1b.dat, E1c.dat (memory codes, primary chips)
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define GC_HD __host__ __device__
#define GC_TABLE_DEF(type, name) __device__ type name##_dev
#else
#define GC_HD
#endif

namespace gc {
namespace codegen {

GC_HD inline int parity32(uint32_t v)
{
    v ^= v >> 16; v ^= v >> 8; v ^= v >> 4; v ^= v >> 2; v ^= v >> 1;
    return (int)(v & 1u);
}

// ---- GPS L5: stage i of a 13-stage register is bit i-1; the registers shift towards stage 13, whose content is the output ----
GC_HD inline uint32_t l5_taps(const int* pos, int n)
{
    uint32_t m = 0;
    for (int i = 0; i < n; ++i) m |= 1u << (pos[i] - 1);
    return m;
}
GC_HD inline void gen_l5(int advance, int8_t* out, int n = 10230)
{
    const int xaPos[4] = {9, 10, 12, 13}, xbPos[8] = {1, 3, 4, 6, 7, 8, 12, 13};
    const uint32_t ta = l5_taps(xaPos, 4), tb = l5_taps(xbPos, 8), all = 0x1FFFu;
    const uint32_t resetState = all & ~(1u << 11);             // stages 1..11 and 13 set, stage 12 clear (generateL5Icode.m:53)
    uint32_t xa = all, xb = all;
    for (int i = 0; i < advance; ++i) xb = ((xb << 1) & all) | (uint32_t)parity32(xb & tb);   // :112-118
    for (int i = 0; i < n; ++i) {
        const int a = (xa >> 12) & 1, b = (xb >> 12) & 1;
        out[i] = (int8_t)(1 - 2 * (a ^ b));                    // XBI .* XA (:132)
        xa = (xa == resetState) ? all : (((xa << 1) & all) | (uint32_t)parity32(xa & ta));     // :57-66
        xb = ((xb << 1) & all) | (uint32_t)parity32(xb & tb);
    }
}

// ---- Galileo E5: 14-stage registers written as the binary number the reference's digit vectors spell (element 1 = bit 13 = the
//      output); feedback = parity of the tapped stages, shifted in at element 14 ----
GC_HD inline void gen_gal_e5(int startValue, int fbOctal1, int fbOctal2, int8_t* out)
{
    const uint32_t t1 = (uint32_t)fbOctal1 >> 1, t2 = (uint32_t)fbOctal2 >> 1, all = 0x3FFFu;   // dec2bin(.)(1:14): the last digit is dropped
    uint32_t r1 = all, r2 = (uint32_t)startValue & all;
    for (int i = 0; i < 10230; ++i) {
        const int o1 = ((r1 >> 13) & 1) & (int)((t1 >> 13) & 1), o2 = ((r2 >> 13) & 1) & (int)((t2 >> 13) & 1);   // RegOut(1) = Register(1) * taps(1)
        out[i] = (int8_t)((1 - 2 * o1) * (1 - 2 * o2));
        const uint32_t f1 = (uint32_t)parity32(r1 & t1), f2 = (uint32_t)parity32(r2 & t2);
        r1 = ((r1 << 1) & all) | f1;
        r2 = ((r2 << 1) & all) | f2;
    }
}
GC_HD inline int hexval(char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c - 'A' + 10; }
// 25 hex characters -> 100 chips (the 13-character and the 12-character half are each a left-padded binary number)
GC_HD inline void gen_gal_secondary(const char* hex25, int8_t* out)
{
    for (int i = 0; i < 25; ++i) {
        const int v = hexval(hex25[i]);
        for (int b = 0; b < 4; ++b) out[4 * i + b] = (int8_t)(1 - 2 * ((v >> (3 - b)) & 1));
    }
}

// ---- BDS B2a: 13-stage registers, element 1 = bit 12, the registers shift towards element 13 (bit 0, the output) ----
GC_HD inline uint32_t b2a_taps(const int* pos, int n)
{
    uint32_t m = 0;
    for (int i = 0; i < n; ++i) m |= 1u << (13 - pos[i]);
    return m;
}
GC_HD inline void gen_b2a(int reg2Init, int pilot, int8_t* out, int n = 10230)
{
    const int d1[4] = {1, 5, 11, 13}, d2[6] = {3, 5, 9, 11, 12, 13}, p1[4] = {3, 6, 7, 13}, p2[6] = {1, 5, 7, 8, 12, 13};
    const uint32_t t1 = pilot ? b2a_taps(p1, 4) : b2a_taps(d1, 4), t2 = pilot ? b2a_taps(p2, 6) : b2a_taps(d2, 6), all = 0x1FFFu;
    uint32_t r1 = all, r2 = (uint32_t)reg2Init & all;
    for (int i = 1; i <= n; ++i) {
        out[i - 1] = (int8_t)(1 - 2 * (int)((r1 ^ r2) & 1u));  // register1(end) * register2(end)
        r1 = (r1 >> 1) | ((uint32_t)parity32(r1 & t1) << 12);
        r2 = (r2 >> 1) | ((uint32_t)parity32(r2 & t2) << 12);
        if (i == 8190) r1 = all;                               // generateB2aDataCode.m:135-137
    }
}

// ---- BDS B1I: 11-stage G1 / G2, stage i = bit i-1, initial phase 01010101010, shifting towards stage 11 ----
GC_HD inline void gen_b1i(int s1, int s2, int s3 /* 0 = two taps */, int8_t* out)
{
    const uint32_t all = 0x7FFu;
    uint32_t init = 0;
    for (int i = 1; i <= 11; ++i) init |= (uint32_t)((i % 2) == 0) << (i - 1);   // -1*[-1 1 -1 ...]: stages 2, 4, ... hold logic 1
    uint32_t g1 = init, g2 = init;
    const uint32_t t1 = (1u << 0) | (1u << 6) | (1u << 7) | (1u << 8) | (1u << 9) | (1u << 10);                       // 1 7 8 9 10 11
    const uint32_t t2 = (1u << 0) | (1u << 1) | (1u << 2) | (1u << 3) | (1u << 4) | (1u << 7) | (1u << 8) | (1u << 10); // 1 2 3 4 5 8 9 11
    const uint32_t sel = (1u << (s1 - 1)) ^ (1u << (s2 - 1)) ^ (s3 ? (1u << (s3 - 1)) : 0u);
    for (int i = 0; i < 2046; ++i) {
        const int a = (g1 >> 10) & 1, b = parity32(g2 & sel);
        out[i] = (int8_t)(-(1 - 2 * (a ^ b)));                 // CAcode = -(g1 .* g2)  (:102)
        g1 = ((g1 << 1) & all) | (uint32_t)parity32(g1 & t1);
        g2 = ((g2 << 1) & all) | (uint32_t)parity32(g2 & t2);
    }
}

// ---- GPS L2C: 27-stage register in the number its initial state spells (element 1 = bit 26, element 27 = bit 0 = the output); the
//      output re-enters at element 1 and is added to the tapped elements.  `phase` 0: CM  [c 0 c 0 ...], 1: CL  [0 c 0 c ...] ----
GC_HD inline void gen_l2c(uint32_t init, long long nChips, int phase, int8_t* out)
{
    const int pos[11] = {4, 7, 9, 12, 15, 17, 19, 22, 23, 24, 25};
    uint32_t mask = 0;
    for (int i = 0; i < 11; ++i) mask |= 1u << (27 - pos[i]);
    uint32_t r = init & 0x7FFFFFFu;
    for (long long i = 0; i < nChips; ++i) {
        const uint32_t o = r & 1u;
        out[2 * i + phase] = (int8_t)(1 - 2 * (int)o);
        out[2 * i + 1 - phase] = 0;
        r = (r >> 1) | (o << 26);
        if (o) r ^= mask;
    }
}

// ---- BDS B1C: Weil code from the Legendre sequence of N = 10243; `leg` = N bytes of scratch for that sequence ----
GC_HD inline void legendre_sequence(int N, uint8_t* leg)
{
    for (int i = 0; i < N; ++i) leg[i] = 0;
    for (long long x = 1; x < N; ++x) leg[(x * x) % N] = 1;   // the quadratic residues: JacobiSymbol(k, N) == 1
}
// mode 0: data BOC(1,1) [-c c], 1: pilot BOC(1,1), 2: pilot BOC(6,1) (twelve entries (-1)^ii * c per chip)
GC_HD inline void gen_b1c(int w, int p, int mode, const uint8_t* leg, int8_t* out)
{
    const int N = 10243;
    for (int ind = 0; ind < 10230; ++ind) {
        const int k = (ind + p - 1) % N;
        const int c = 1 - 2 * (int)(leg[k] ^ leg[(k + w) % N]);
        if (mode == 2) {
            for (int ii = 1; ii <= 12; ++ii) out[12 * ind + ii - 1] = (int8_t)((ii & 1) ? -c : c);
        } else {
            out[2 * ind] = (int8_t)-c;
            out[2 * ind + 1] = (int8_t)c;
        }
    }
}

// ---- Galileo E1 memory codes: 1023 hex characters -> 4092 primary chips ----
GC_HD inline void gen_e1(const char* hex1023, int8_t* out)
{
    for (int i = 0; i < 1023; ++i) {
        const int v = hexval(hex1023[i]);
        for (int b = 0; b < 4; ++b) out[4 * i + b] = (int8_t)(1 - 2 * ((v >> (3 - b)) & 1));
    }
}

}  // namespace codegen
}  // namespace gc
