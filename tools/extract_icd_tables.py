#!/usr/bin/env python3
"""Extracts the ICD constant tables that the reference's code generators embed (or read from E1b.dat / E1c.dat at run time) and
writes them, as data only, to

    cu-sdr-collection_b200/csrc/icd_tables.inc   (C arrays for the library's generators, csrc/codegen.h)
    oracle/icd_tables.py                          (the same numbers for the oracle's restatement of the generators)

Run in the authoring container, where /root/reference is mounted; the outputs are committed (the GPU box has no reference tree).
Sources (all under /root/reference): GPS/GPS_L5C/include/generateL5Icode.m:62-89, generateL5Qcode.m (XB code advances);
GAL/GAL_E5a/include/generateE5aIcode.m / generateE5aQcode.m / generateE5aQ_secondary.m, GAL/GAL_E5b/include/generateE5bIcode.m /
generateE5bQcode.m / generateE5bQ_secondary.m (start values, secondary codes); BDS/B2a/include/generateB2aDataCode.m /
generateB2aPilotCode.m (register-2 initial states); BDS/B1C/include/generateDataBOC11.m / generatePilotBOC11.m (Weil w, p);
BDS/B1I/include/generateCAcode53.m (G2 phase selections); GPS/GPS_L2C/include/generateCMcode.m / generateCLcode.m (initial states);
GAL/GAL_E1C/include/E1b.dat / E1c.dat (memory codes)."""
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def src(path):
    """File text without MATLAB comments (full-line % and the '... text' continuation comments inside tables)."""
    out = []
    for line in open(os.path.join(REF, path), errors="ignore"):
        s = line.strip()
        if s.startswith("%"):
            continue
        if s.startswith("..."):
            continue
        line = re.sub(r"\.\.\..*$", "", line)
        out.append(line)
    return "".join(out)


def block(text, name):
    """The bracketed literal assigned to `name`."""
    m = re.search(re.escape(name) + r"\s*=\s*\[(.*?)\]\s*;", text, re.S)
    assert m, name
    return m.group(1)


def ints(text):
    return [int(x) for x in re.findall(r"-?\d+", text)]


def quoted(text):
    return re.findall(r"'([0-9A-Fa-f]+)'", text)


T = {}
T["L5I_ADVANCE"] = ints(block(src("GPS/GPS_L5C/include/generateL5Icode.m"), "l5i_init"))
T["L5Q_ADVANCE"] = ints(block(src("GPS/GPS_L5C/include/generateL5Qcode.m"), "l5q_init"))
T["E5AI_START"] = [int(x, 8) for x in quoted(block(src("GAL/GAL_E5a/include/generateE5aIcode.m"), "e5ai_init"))]
T["E5AQ_START"] = [int(x, 8) for x in quoted(block(src("GAL/GAL_E5a/include/generateE5aQcode.m"), "e5aq_init"))]
T["E5BI_START"] = [int(x, 8) for x in quoted(block(src("GAL/GAL_E5b/include/generateE5bIcode.m"), "e5bi_init2"))]
T["E5BQ_START"] = [int(x, 8) for x in quoted(block(src("GAL/GAL_E5b/include/generateE5bQcode.m"), "e5bq_init2"))]
E5AQ_SEC = quoted(block(src("GAL/GAL_E5a/include/generateE5aQ_secondary.m"), "secondary_code"))
E5BQ_SEC = quoted(block(src("GAL/GAL_E5b/include/generateE5bQ_secondary.m"), "secondary_code"))
for k in ("E5AI_START", "E5AQ_START", "E5BI_START", "E5BQ_START"):
    assert len(T[k]) == 50, (k, len(T[k]))
assert len(E5AQ_SEC) == 50 and len(E5BQ_SEC) == 50 and all(len(x) == 25 for x in E5AQ_SEC + E5BQ_SEC)


def b2a_states(path):
    rows = re.findall(r"((?:[01]\s+){12}[01])", block(src(path), "B2aData_reg2_ini"))
    assert len(rows) == 63, len(rows)
    return [int("".join(r.split()), 2) for r in rows]          # bit 12 = register element 1 ... bit 0 = element 13


T["B2AD_REG2"] = b2a_states("BDS/B2a/include/generateB2aDataCode.m")
T["B2AP_REG2"] = b2a_states("BDS/B2a/include/generateB2aPilotCode.m")
wd = ints(block(src("BDS/B1C/include/generateDataBOC11.m"), "wp_data"))
wp = ints(block(src("BDS/B1C/include/generatePilotBOC11.m"), "wp_pilot"))
assert len(wd) == 126 and len(wp) == 126
T["B1CD_W"], T["B1CD_P"] = wd[0::2], wd[1::2]
T["B1CP_W"], T["B1CP_P"] = wp[0::2], wp[1::2]
assert ints(block(src("BDS/B1C/include/generatePilotBOC61.m"), "wp_pilot")) == wp
# B1I: the G2 phase selections are written with repmat([..], 1, n)
b1i = src("BDS/B1I/include/generateCAcode53.m")


def b1i_sel(name):
    txt = block(b1i, name)
    txt = re.sub(r"repmat\(\[(\d+)\]\s*,\s*1\s*,\s*(\d+)\)", lambda m: ", ".join([m.group(1)] * int(m.group(2))), txt)
    return ints(txt)


T["B1I_G2S1"], T["B1I_G2S2"], T["B1I_G2S3"] = b1i_sel("g2s1"), b1i_sel("g2s2"), b1i_sel("g2s3")
assert len(T["B1I_G2S1"]) == 58 and len(T["B1I_G2S2"]) == 58 and len(T["B1I_G2S3"]) == 21, [len(T[k]) for k in ("B1I_G2S1", "B1I_G2S2", "B1I_G2S3")]
# L2C: the literals are octal digit strings written as decimal numbers (leading zeros dropped by MATLAB: same octal value)
T["L2CM_INIT"] = [int(str(x), 8) for x in ints(block(src("GPS/GPS_L2C/include/generateCMcode.m"), "l2cm_init"))]
T["L2CL_INIT"] = [int(str(x), 8) for x in ints(block(src("GPS/GPS_L2C/include/generateCLcode.m"), "l2cl_init"))]
assert len(T["L2CM_INIT"]) == 115 and len(T["L2CL_INIT"]) == 115, (len(T["L2CM_INIT"]), len(T["L2CL_INIT"]))
assert len(T["L5I_ADVANCE"]) == 210 and len(T["L5Q_ADVANCE"]) == 210


def memory_codes(name):
    v = open(os.path.join(REF, "GAL/GAL_E1C/include", name)).read().split()
    assert len(v) >= 50 * 4092
    rows = []
    for prn in range(50):
        bits = "".join(v[prn * 4092:(prn + 1) * 4092])
        rows.append("%0*X" % (1023, int(bits, 2)))
    return rows


E1B, E1C = memory_codes("E1b.dat"), memory_codes("E1c.dat")

hdr = ("ICD constant tables of the code generators, extracted from the reference tree by tools/extract_icd_tables.py "
       "(data only; see that script for the source of every table)")
with open(os.path.join(ROOT, "oracle", "icd_tables.py"), "w") as f:
    f.write('"""%s."""\n' % hdr)
    for k, v in T.items():
        f.write("%s = %r\n" % (k, v))
    for k, v in (("E5AQ_SECONDARY", E5AQ_SEC), ("E5BQ_SECONDARY", E5BQ_SEC), ("E1B_HEX", E1B), ("E1C_HEX", E1C)):
        f.write("%s = [\n" % k)
        for x in v:
            f.write("    %r,\n" % x)
        f.write("]\n")
with open(os.path.join(ROOT, "cu-sdr-collection_b200", "csrc", "icd_tables.inc"), "w") as f:
    f.write("// %s.\n" % hdr)
    for k, v in T.items():
        f.write("static const int k%s[%d] = {%s};\n" % (k, len(v), ", ".join(str(x) for x in v)))
    for k, v in (("E5AQ_SECONDARY", E5AQ_SEC), ("E5BQ_SECONDARY", E5BQ_SEC), ("E1B_HEX", E1B), ("E1C_HEX", E1C)):
        f.write("static const char* const k%s[%d] = {\n" % (k, len(v)))
        for x in v:
            f.write('    "%s",\n' % x)
        f.write("};\n")
print({k: len(v) for k, v in T.items()}, len(E1B), len(E1B[0]))
