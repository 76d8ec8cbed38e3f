"""Extracts the known-answer vectors of g2o's linear-solver unit test into a small JSON fixture.

Source (read-only, only available in the authoring container):
  /root/reference/third_party/g2o/unit_test/solver/sparse_system_helper.cpp
    :52-153  36x36 block-sparse SPD matrix (12 blocks of 3), upper triangle, "BLOCK : r c" + 3 rows
    :155-194 its dense inverse
    :255-296 right-hand side b      :298-340 solution x
  asserted by unit_test/solver/linear_solver_test.cpp:73-87 with isApprox(1e-6).
Only the numbers are extracted (test data), no code. Run once: `python tests/golden/make_g2o_kat.py`.
"""
import json
import os
import re

SRC = "/root/reference/third_party/g2o/unit_test/solver/sparse_system_helper.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "g2o_linear_solver_kat.json")


def main():
    txt = open(SRC).read()
    # --- sparse matrix string
    body = txt[txt.index("std::string sparseMatrixString()"):txt.index("std::string denseInverseMatrixString()")]
    lines = re.findall(r'aux << "([^"]*)"', body)
    rbi = [int(t) for t in lines[0].split(":")[1].split()][1:]
    blocks = []
    i = 2
    while i < len(lines):
        m = re.match(r"BLOCK : (\d+) (\d+)", lines[i])
        assert m, lines[i]
        r, c = int(m.group(1)), int(m.group(2))
        rows = [[float(v) for v in lines[i + 1 + k].split()] for k in range(3)]
        blocks.append(dict(r=r, c=c, m=rows))
        i += 4
    # --- dense inverse
    body = txt[txt.index("std::string denseInverseMatrixString()"):txt.index("createTestVectorB()")]
    lines = re.findall(r'aux << "([^"#]*)"', body)
    inv = [[float(v) for v in ln.split()] for ln in lines if ln.strip()]
    assert len(inv) == 36 and all(len(r) == 36 for r in inv)

    def vec(name, end):
        b = txt[txt.index(name):]
        b = b[:b.index(end)]
        return [float(v) for v in re.findall(r"result\(idx\+\+\) = ([-0-9.e+]+);", b)]

    bvec = vec("g2o::VectorX createTestVectorB()", "return result;")
    xvec = vec("g2o::VectorX createTestVectorX()", "return result;")
    assert len(bvec) == 36 and len(xvec) == 36
    json.dump(dict(source="third_party/g2o/unit_test/solver/sparse_system_helper.cpp", block_ends=rbi, blocks=blocks,
                   inverse=inv, b=bvec, x=xvec, tolerance=1e-6), open(OUT, "w"))
    print("wrote", OUT, len(blocks), "blocks")


if __name__ == "__main__":
    main()
