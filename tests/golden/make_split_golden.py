"""Extracts input / expected prototxt pairs of the reference's own InsertSplits tests (src/caffe/test/test_split_layer.cpp:
SplitLayerInsertionTest.TestInsertion :688, .TestInsertionTwoTop :783, .TestWithInPlace :889) into
tests/golden/insert_splits_fixtures.json; oracle.ref.insert_splits (what drives the reference's layers as a net that can run
Backward) is checked against them.  Run in the container that has /root/reference."""
import json
import os
import re

SRC = "/root/reference/src/caffe/test/test_split_layer.cpp"
HERE = os.path.dirname(os.path.abspath(__file__))


def literal(text, var, start):
    i = text.index("const string& %s =" % var, start)
    j = text.index(";", i)
    return "".join(m.group(1) for m in re.finditer(r'"((?:[^"\\]|\\.)*)"', text[i:j])), j


def main():
    text = open(SRC).read()
    cases = []
    for test in ("TestInsertion)", "TestInsertionTwoTop)", "TestWithInPlace)"):
        pos = text.index("TEST_F(SplitLayerInsertionTest, " + test)
        a, pos = literal(text, "input_proto", pos)
        b, pos = literal(text, "expected_output_proto", pos)
        cases.append({"test": test[:-1], "input": a, "expected": b})
    with open(os.path.join(HERE, "insert_splits_fixtures.json"), "w") as f:
        json.dump({"source": "src/caffe/test/test_split_layer.cpp SplitLayerInsertionTest", "cases": cases}, f, indent=1)
    print("wrote", [c["test"] for c in cases])


if __name__ == "__main__":
    main()
