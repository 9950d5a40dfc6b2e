"""Extracts the V1 -> V2 prototxt pairs of the reference's own upgrade tests (src/caffe/test/test_upgrade_proto.cpp,
NetUpgradeTest.TestSimple :1126-1350 and .TestImageNet :1853-2891: `expected_v1_proto` must upgrade to `expected_v2_proto`) into
tests/golden/upgrade_v1_fixtures.json.  Run in the container that has /root/reference."""
import json
import os
import re

SRC = "/root/reference/src/caffe/test/test_upgrade_proto.cpp"
HERE = os.path.dirname(os.path.abspath(__file__))


def literals(text, var, start):
    i = text.index("const string& %s =" % var, start)
    j = text.index(";", i)
    body = text[i:j]
    return "".join(m.group(1) for m in re.finditer(r'"((?:[^"\\]|\\.)*)"', body)).replace('\\"', '"'), j


def main():
    text = open(SRC).read()
    out, pos = [], text.index("TEST_F(NetUpgradeTest, TestSimple)")
    for _ in range(2):
        v1, pos = literals(text, "expected_v1_proto", pos)
        v2, pos = literals(text, "expected_v2_proto", pos)
        out.append({"v1": v1, "v2": v2})
        nxt = text.find("TEST_F(NetUpgradeTest, TestImageNet)", pos)
        if nxt < 0:
            break
        pos = nxt
    with open(os.path.join(HERE, "upgrade_v1_fixtures.json"), "w") as f:
        json.dump({"source": "src/caffe/test/test_upgrade_proto.cpp NetUpgradeTest.TestSimple / TestImageNet", "cases": out}, f, indent=1)
    print("wrote", len(out), "cases", [len(c["v1"]) for c in out])


if __name__ == "__main__":
    main()
