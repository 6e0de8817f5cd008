"""TEST INFRASTRUCTURE (container-only): assert that recmv_b200/csrc/mc_tables.h carries the same
case tables as the reference's MCGpu/CudaKernels.cu:4-302, and that the edge-flag table the kernels
derive (an edge is cut iff its two corners differ in sign) equals the reference's aiCubeEdgeFlags.
Run: python oracle/check_mc_tables.py
"""
import os
import re
import sys

REF = os.environ.get("RECMV_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def ints(text):
    return [int(t, 0) for t in re.findall(r"-?(?:0x[0-9a-fA-F]+|\d+)", text)]


def block(src, name):
    i = src.index(name)
    j = src.index("{", src.index("=", i))
    depth, k = 0, j
    while True:
        if src[k] == "{":
            depth += 1
        elif src[k] == "}":
            depth -= 1
            if depth == 0:
                break
        k += 1
    return src[j:k + 1]


def main():
    ref = open(os.path.join(REF, "MCGpu/CudaKernels.cu")).read()
    mine = open(os.path.join(HERE, "../recmv_b200/csrc/mc_tables.h")).read()
    ref_tri = ints(block(ref, "a2iTriangleConnectionTable[256][16]"))
    # our header keeps the table in compact form: one hex digit (edge id) per triangle corner, one string per case
    strs = re.findall(r'"([0-9a-b]*)"', block(mine, "kMcTriHex[256]"))
    assert len(strs) == 256, len(strs)
    my_tri = []
    for h in strs:
        row = [int(ch, 16) for ch in h]
        my_tri += row + [-1] * (16 - len(row))
    assert len(ref_tri) == len(my_tri) == 4096, (len(ref_tri), len(my_tri))
    bad = [i // 16 for i in range(4096) if ref_tri[i] != my_tri[i]]
    assert not bad, f"tri table differs in cases {sorted(set(bad))}"
    ref_flags = ints(block(ref, "aiCubeEdgeFlags[256]"))
    edges = ints(block(mine, "kMcEdgeCorners[12][2]"))
    for case in range(256):
        f = 0
        for e in range(12):
            a, b = edges[2 * e], edges[2 * e + 1]
            if ((case >> a) & 1) != ((case >> b) & 1):
                f |= 1 << e
        assert f == ref_flags[case], (case, hex(f), hex(ref_flags[case]))
        used = {t for t in my_tri[case * 16:(case + 1) * 16] if t >= 0}
        assert used == {e for e in range(12) if f >> e & 1}, case
    ref_conn = ints(block(ref, "a2iEdgeConnection[12][2]"))
    assert ref_conn == edges
    print("mc tables identical to reference; edge flags derivable; every cut edge is used")


if __name__ == "__main__":
    sys.exit(main())
