"""TEST INFRASTRUCTURE.  Writes patched copies of the reference's db_construction.cpp and rna_interaction_search.cpp
into the (git-ignored) build directory: the three functions INTEGRATION.md replaces are swapped for
oracle/integration/*.inc, nothing else changes.

  python apply_patch.py /root/reference/src oracle/_ref/patched
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def replace_function(src: str, head: str, body: str) -> str:
    a = src.index(head)
    i = src.index("{", a)
    depth = 0
    while True:
        if src[i] == "{":
            depth += 1
        elif src[i] == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return src[:a] + body.rstrip() + "\n" + src[i + 1:]


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    inc = lambda name: open(os.path.join(HERE, name)).read()
    head = '#include "priblast_acc.h"\n#include <cstdlib>\n#include <fstream>\n#include <mutex>\n'
    src = open(os.path.join(ref, "db_construction.cpp")).read()
    src = replace_function(src, "void DbConstruction::CalculateAccessibility(", inc("CalculateAccessibility.inc"))
    src = replace_function(src, "void DbConstruction::ConstructSuffixArray(", inc("ConstructSuffixArray.inc"))
    open(os.path.join(out, "db_construction.cpp"), "w").write(head + src)
    src = open(os.path.join(ref, "rna_interaction_search.cpp")).read()
    src = replace_function(src, "void RnaInteractionSearch::CalculateAccessibility(", inc("RisCalculateAccessibility.inc"))
    open(os.path.join(out, "rna_interaction_search.cpp"), "w").write(head + src)


if __name__ == "__main__":
    main()
