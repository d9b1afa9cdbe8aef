"""One-off validation (runs only where /root/reference is mounted): the oracle regenerates the
reference's pre-baked fixed Huffman tables (src/inftree.ts:19-63) with its huft_build port;
compare them entry by entry with the literals in the reference source.  Nothing is copied."""
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/src/inftree.ts"


def ref_array(src, name):
    m = re.search(r"const %s = \[(.*?)\];" % name, src, re.S)
    return [int(x) for x in re.findall(r"-?\d+", m.group(1))]


def main():
    if not os.path.exists(REF):
        print("reference not mounted; skipped")
        return 0
    src = open(REF).read()
    tl, td = O.fixed_tables()
    ok = tl == ref_array(src, "fixed_tl") and td == ref_array(src, "fixed_td")
    print("fixed_tl entries:", len(tl) // 3, "fixed_td entries:", len(td) // 3, "identical to reference:", ok)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
